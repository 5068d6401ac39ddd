"""Feature-sharded TopK SAE training (d=1280, n=81920 over 8 GPUs).

Rank r owns dictionary rows [r*n/G, (r+1)*n/G) of W_enc, b_enc and W_dec plus their Adam state; x and b_dec are
replicated.  One step (equal to the single-GPU step of train_sae.py:421-453 on the same batch):

  local fused encode (top-32 of the shard)          -- freud_topk_encode
  all-gather the G candidate lists, global top-32   -- NCCL all_gather + freud_shard_merge
  decode only the winners this rank owns            -- freud_shard_localize + freud_topk_decode (partial sums)
  all-reduce the partial reconstructions [N,d]      -- NCCL all_reduce          (the one large exchange)
  residual / SSE / losses                           -- freud_residual + freud_topk_loss_scalars
  backward is rank-local (e is replicated): dacts, CSC of the owned winners, dW_dec / dW_enc / db_enc
  db_dec needs one [d] all-reduce of -db_enc^T W_enc; the clip norm one scalar all-reduce; Adam is local.

AuxK (topkautoencoder.py:109-127) follows the same pattern once a latent can be dead: every rank selects the top
k_aux = min(d/2, num_dead) pre-activations among ITS dead latents, the candidate lists are all-gathered, the global
top-k_aux is taken from them (value descending, dictionary index ascending -- rank-major concatenation keeps that
order), each rank decodes the winners it owns, and the partial e_hat is all-reduced.  Its backward is rank-local like
the main one.  multi-TopK (topkautoencoder.py:129-138: a second decode of the top 4k latents, its FVU weighted 1/8) is
selected the same way over ALL latents of every shard (local top-4k of the materialised shard pre-activations ->
all-gather -> global top-4k), decoded in parts and all-reduced; as in the reference, the returned reconstruction /
activations / indices and the did_fire bookkeeping are then those of the 4k selection.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import ops, topk_engine
from ._lib import BF16, FP32
from .optim import FusedAdam
from .trainer import build_scheduler

_KEYS = ("encoder.weight", "encoder.bias", "W_dec", "b_dec")


class FeatureShardedTopKTrainer:
    def __init__(self, full_state: dict, k: int, *, lr, steps, clip_thresh=1.0, scheduler="linear",
                 scheduler_params=None, precision="bf16", device=None, group=None, auxk_alpha=0.0,
                 dead_feature_threshold=None, multi_topk=False):
        """full_state: the reference state_dict (W_dec, b_dec, encoder.weight, encoder.bias) -- every rank slices
        its own rows, so a checkpoint written by the single-GPU model loads unchanged."""
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        if k != ops.K_FUSED:
            raise NotImplementedError("feature-sharded mode uses the fused k == 32 encoder")
        self.group = group
        self.G = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        n, d = full_state["encoder.weight"].shape
        if n % self.G:
            raise ValueError("dictionary size must divide evenly over the ranks")
        self.n, self.d, self.k = n, d, k
        self.n_local = n // self.G
        self.lo = self.rank * self.n_local
        dev = torch.device(device if device is not None else torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("feature-sharded training runs on CUDA only (no CPU fallback)")
        sl = slice(self.lo, self.lo + self.n_local)
        self.params = {
            "encoder.weight": full_state["encoder.weight"][sl].detach().to(dev, torch.float32).contiguous(),
            "encoder.bias": full_state["encoder.bias"][sl].detach().to(dev, torch.float32).contiguous(),
            "W_dec": full_state["W_dec"][sl].detach().to(dev, torch.float32).contiguous(),
            "b_dec": full_state["b_dec"].detach().to(dev, torch.float32).contiguous(),
        }
        self.plist = [torch.nn.Parameter(self.params[k_]) for k_ in _KEYS]
        for p in self.plist:
            p.grad = torch.zeros_like(p)
        self.precision = {"bf16": BF16, "fp32": FP32}[precision]
        self.clip_thresh = clip_thresh
        self.optimizer = FusedAdam(self.plist, lr=lr, max_grad_norm=clip_thresh)
        self.scheduler = build_scheduler(self.optimizer, scheduler, scheduler_params or {}, steps)
        self.num_frames_since_fired = torch.zeros(self.n_local, device=dev, dtype=torch.long)
        self.device = dev
        self.auxk_alpha = float(auxk_alpha)
        self.multi_topk = bool(multi_topk)
        self.dead_feature_threshold = dead_feature_threshold
        self.tokens_seen = 0

    def _allreduce(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def gathered_state(self) -> dict:
        """Full (unsharded) state_dict on every rank, in the reference layout."""
        out = {}
        for key, p in zip(_KEYS, self.plist):
            if key == "b_dec":
                out[key] = p.data.clone()
                continue
            parts = [torch.empty_like(p.data) for _ in range(self.G)]
            dist.all_gather(parts, p.data.contiguous(), group=self.group)
            out[key] = torch.cat(parts, 0)
        return out

    def gather_batch(self, local_files: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """Feed for the replicated input: every rank uploads only its B/G files (its own PCIe link) and the parts
        are all-gathered over NVLink into the full [B,T,d] batch, instead of G full host->device copies."""
        per, T, d = local_files.shape
        if out is None:
            out = torch.empty((per * self.G, T, d), dtype=local_files.dtype, device=local_files.device)
        dist.all_gather_into_tensor(out, local_files.contiguous(), group=self.group)
        return out

    def step(self, x: torch.Tensor):
        if not x.is_cuda:
            raise RuntimeError("activations must already be on the CUDA device")
        x = x.float().contiguous()
        B, T, d = x.shape
        N, k, prec = B * T, self.k, self.precision
        W_enc, b_enc, W_dec, b_dec = (p.data for p in self.plist)
        x2 = x.view(N, d)
        xc_hi, xc_lo, tv = ops.topk_prep_x(x, b_dec, prec)
        we_hi, we_lo = ops.split_operand(W_enc, prec)
        wd = ops.split_operand(W_dec, BF16)[0] if prec == BF16 else W_dec
        lvals, lidx = ops.topk_encode(xc_hi, xc_lo, we_hi, we_lo, b_enc, prec)          # shard-local top-32
        all_vals = torch.empty((self.G, N, k), dtype=torch.float32, device=x.device)
        all_idx = torch.empty((self.G, N, k), dtype=torch.int32, device=x.device)
        dist.all_gather_into_tensor(all_vals, lvals, group=self.group)
        dist.all_gather_into_tensor(all_idx, lidx, group=self.group)
        top_vals, top_gidx = ops.shard_merge(all_vals, all_idx, self.n_local)           # identical on every rank
        own_vals, own_idx = ops.shard_localize(top_vals, top_gidx, self.lo, self.n_local)
        bias = b_dec if self.rank == 0 else torch.zeros_like(b_dec)                     # b_dec enters the sum once
        partial, _, _, _ = ops.topk_decode(own_vals, own_idx, wd, bias)
        sae_out = self._allreduce(partial)                                              # [N,d] fp32
        # a latent can only be dead once more than `thr` frames have been seen (same short-cut as SAETrainer)
        thr = self.dead_feature_threshold
        num_dead, dead_local = 0, None
        if thr is not None and self.auxk_alpha != 0.0 and self.tokens_seen > thr:
            dead_local = self.num_frames_since_fired > thr
            num_dead = int(self._allreduce(dead_local.sum()))  # the reference's own read-back (:109), over all shards
        g = {k_: p.grad for k_, p in zip(_KEYS, self.plist)}
        xc = xc_hi if prec == BF16 else x2
        auxk = torch.zeros((), dtype=torch.float32, device=x.device)
        mfvu = torch.zeros((), dtype=torch.float32, device=x.device)
        ret_vals, ret_gidx, ret_out = top_vals, top_gidx, sae_out
        if num_dead == 0 and not self.multi_topk:
            rdt = torch.bfloat16 if prec == BF16 else torch.float32
            e, sse, colsum_e = ops.residual(sae_out, x2, rdt)
            scal = ops.topk_loss_scalars(sse, tv, N * d)
            # ---- backward (rank-local): loss = fvu
            scales = scal[2:4]
            dacts = ops.topk_dacts(e, own_idx, wd)
            offsets, entries = ops.csc_build(own_idx, self.n_local)
            ops.topk_sparse_grads(offsets, entries, own_vals, dacts, e, xc, b_dec, scales, g["W_dec"],
                                  g["encoder.weight"], g["encoder.bias"], k, False)
            ops.topk_bdec_grad(colsum_e if self.rank == 0 else None, scales if self.rank == 0 else None,
                               g["encoder.bias"], we_hi if prec == BF16 else W_enc, g["b_dec"], False)
        else:
            e, sse, colsum_e = ops.residual(sae_out, x2, torch.float32)
            scal = ops.topk_loss_scalars(sse, tv, N * d)
            # ---- backward (rank-local) through the generic engine path: one decode per selection (main, aux, 4k) on
            # this shard's rows; replicated terms (the direct b_dec gradient) enter on rank 0 only, the sum over ranks
            # completes them
            st = topk_engine.TopKState(prec, x2, xc_hi, wd, we_hi if prec == BF16 else W_enc, b_dec, k, self.n_local,
                                       scal, True, own_vals,
                                       own_idx, e, colsum_e if self.rank == 0 else torch.zeros_like(colsum_e),
                                       auxk_alpha=self.auxk_alpha)
            if num_dead > 0:
                k_aux = d // 2
                scale = min(num_dead / k_aux, 1.0)
                k_aux = min(k_aux, num_dead)
                a_vals, a_gidx = self._select_over_shards(xc_hi, xc_lo, we_hi, we_lo, b_enc, dead_local, k_aux, N, prec)
                a_own_vals, a_own_idx = ops.shard_localize(a_vals, a_gidx, self.lo, self.n_local)
                partial, _, _, _ = ops.topk_decode(a_own_vals, a_own_idx, wd, bias)
                e_hat = self._allreduce(partial)
                r_aux, sse_aux, _ = ops.residual(e_hat, e, torch.float32, want_colsum=False)  # e_hat - e, e not detached
                auxk = (scale * sse_aux[0] / scal[4].double()).float() * self.auxk_alpha
                st.aux = (a_own_vals, a_own_idx, r_aux, scale)
            if self.multi_topk:
                m_vals, m_gidx = self._select_over_shards(xc_hi, xc_lo, we_hi, we_lo, b_enc, None, 4 * k, N, prec)
                m_own_vals, m_own_idx = ops.shard_localize(m_vals, m_gidx, self.lo, self.n_local)
                partial, _, _, _ = ops.topk_decode(m_own_vals, m_own_idx, wd, bias)
                m_out = self._allreduce(partial)
                r_m, sse_m, colsum_m = ops.residual(m_out, x2, torch.float32)
                mfvu = (sse_m[0] / scal[4].double()).float()
                st.multi = (m_own_vals, m_own_idx, r_m, colsum_m if self.rank == 0 else torch.zeros_like(colsum_m))
                ret_vals, ret_gidx, ret_out = m_vals, m_gidx, m_out  # the reference rebinds its outputs (:133-138)
            topk_engine.topk_backward(st, 1.0, 1.0, 1.0 / 8.0, out=g)
            offsets = st.offsets  # of the returned encoding
        self._allreduce(g["b_dec"])                                                     # [d]
        # ---- global-norm clip + Adam: b_dec's gradient is replicated, count it once
        tl_local = ops.make_tensor_list([p.data for p in self.plist[:3]], [p.grad for p in self.plist[:3]])
        sumsq = ops.grad_sumsq(tl_local, x.device)
        if self.rank == 0:
            tl_b = ops.make_tensor_list([self.plist[3].data], [self.plist[3].grad])
            sumsq = sumsq + ops.grad_sumsq(tl_b, x.device)
        self._allreduce(sumsq)
        self.optimizer.step(grad_sumsq=sumsq)
        self.scheduler.step()
        ops.dead_latent_update(offsets, self.num_frames_since_fired, N)
        self.tokens_seen += N
        return {"loss": scal[0] + auxk + mfvu / 8, "fvu": scal[0], "auxk_loss": auxk, "multi_topk_fvu": mfvu,
                "grad_sumsq": sumsq, "top_idx": ret_gidx, "top_acts": ret_vals, "sae_out": ret_out}

    def _select_over_shards(self, xc_hi, xc_lo, we_hi, we_lo, b_enc, subset, k_aux, N, prec):
        """Global top-k_aux pre-activations over the latents of all shards -- among this rank's `subset` (bool mask over
        its rows: the dead latents, AuxK) or among all of them (None: multi-TopK): (vals [N,k_aux], global idx)."""
        dev = xc_hi.device
        lv = torch.full((N, k_aux), -1.0, dtype=torch.float32, device=dev)   # absent candidates lose to relu(.) >= 0
        lg = torch.full((N, k_aux), -1, dtype=torch.int32, device=dev)
        if subset is None:
            pre = ops.gemm_nt(xc_hi, xc_lo, we_hi, we_lo, b_enc, True, prec)            # [N, n_local] fp32
            kk = min(k_aux, self.n_local)
            v, loc = ops.row_topk(pre, kk)                                              # (value desc, index asc)
            lv[:, :kk] = v
            lg[:, :kk] = loc + self.lo
            S = 0
        else:
            dead_idx = torch.nonzero(subset).squeeze(1).to(torch.int32)
            S = dead_idx.numel()
        if S > 0:
            ws_hi = ops.gather_rows(we_hi, dead_idx)
            ws_lo = ops.gather_rows(we_lo, dead_idx) if we_lo is not None else None
            bs = ops.gather_rows(b_enc, dead_idx)
            pre_dead = ops.gemm_nt(xc_hi, xc_lo, ws_hi, ws_lo, bs, True, prec)          # [N,S] fp32
            kk = min(k_aux, S)
            v, loc = ops.row_topk(pre_dead, kk)                                         # (value desc, index asc)
            lv[:, :kk] = v
            lg[:, :kk] = ops.index_map(dead_idx, loc) + self.lo
        all_v = torch.empty((self.G, N, k_aux), dtype=torch.float32, device=dev)
        all_g = torch.empty((self.G, N, k_aux), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_v, lv, group=self.group)
        dist.all_gather_into_tensor(all_g, lg, group=self.group)
        # rank-major concatenation: among equal values the lower position is the lower dictionary index
        cat_v = all_v.permute(1, 0, 2).reshape(N, self.G * k_aux).contiguous()
        cat_g = all_g.permute(1, 0, 2).reshape(N, self.G * k_aux).contiguous()
        vals, pos = ops.row_topk(cat_v, k_aux)
        gidx = torch.gather(cat_g, 1, pos.long()).contiguous()
        return vals, gidx

"""CPU restatement of the optimiser side of the train step (train_sae.py:449-451).

TEST INFRASTRUCTURE ONLY.  torch is third-party to the reference (unpinned in
requirements.txt; 2.11.0 here), so these restate torch's published algorithms:
clip_grad_norm_ (torch/nn/utils/clip_grad.py:50,121,165-169), Adam
(torch/optim/adam.py _single_tensor_adam), RAdam (torch/optim/radam.py:256-361)
and the two LR schedules; pinned against the real torch classes in
tests/golden/make_golden.py.
"""
import math

import torch


def clip_grad_norm(grads, max_norm):
    """total = ||[||g_p||_2]_p||_2 ; coef = clamp(max_norm/(total+1e-6), max=1); g *= coef."""
    total = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g) for g in grads]))
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return [g * coef for g in grads], total


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """One Adam update; `step` is the 1-based step count after the increment."""
    m = m + (g - m) * (1 - beta1)  # lerp_
    v = v * beta2 + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def radam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-5, weight_decay=0.0):
    """One RAdam update (non-decoupled L2 weight decay), train_sae.py:375-377."""
    if weight_decay != 0:
        g = g + weight_decay * p
    m = m + (g - m) * (1 - beta1)
    v = v * beta2 + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    m_hat = m / bc1
    rho_inf = 2 / (1 - beta2) - 1
    rho_t = rho_inf - 2 * step * (beta2 ** step) / bc2
    if rho_t > 5.0:
        rect = math.sqrt((rho_t - 4) * (rho_t - 2) * rho_inf / ((rho_inf - 4) * (rho_inf - 2) * rho_t))
        adaptive = math.sqrt(bc2) / (v.sqrt() + eps)
        p = p - m_hat * lr * adaptive * rect
    else:
        p = p - m_hat * lr
    return p, m, v


def cosine_lr(base_lr, step, t_max, eta_min=0.0):
    """Closed form of CosineAnnealingLR (torch/optim/lr_scheduler.py) after `step` scheduler steps."""
    return eta_min + (base_lr - eta_min) * (1 + math.cos(math.pi * step / t_max)) / 2


def linear_warmup_lr(base_lr, step, num_warmup_steps, num_training_steps):
    """transformers/optimization.py:101-131 get_linear_schedule_with_warmup lambda."""
    if step < num_warmup_steps:
        return base_lr * float(step) / float(max(1, num_warmup_steps))
    return base_lr * max(0.0, float(num_training_steps - step) / float(max(1, num_training_steps - num_warmup_steps)))


def dead_latent_update(num_frames_since_fired, top_indices, n_tokens):
    """train_sae.py:442-446: counters += B*T, reset where fired."""
    did_fire = torch.zeros_like(num_frames_since_fired, dtype=torch.bool)
    did_fire[top_indices.flatten()] = True
    out = num_frames_since_fired + n_tokens
    out[did_fire] = 0
    return out

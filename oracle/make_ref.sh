#!/usr/bin/env bash
# Populate oracle/_ref/ with the UNMODIFIED reference (ksadov/FREUD) Python tree so that it travels to the GPU box
# with the repo snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).  The reference is pure Python: there is
# nothing to compile, "building" it is copying the package and its train configs.  Reference sources never enter git.
#
#   bash oracle/make_ref.sh [/root/reference]
#
# Used by: bench.py --impl reference (times the reference's own TopKAutoEncoder + autograd + clip_grad_norm_ +
# torch.optim.Adam step, train_sae.py:421-453), bench.py's parity_check / torch-eager arm, and
# tests/test_gpu_dropin.py (runs the reference's train() on this build's kernels).  Import shims for the two
# packages the image lacks (simple_parsing, whisper) live in oracle/ref_shims.py.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
if [ ! -d "$SRC/src/models" ]; then
  echo "make_ref: $SRC has no src/models (reference tree absent) -- keeping whatever $DST holds" >&2
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/src" "$DST/configs"
for sub in models utils dataset scripts; do
  mkdir -p "$DST/src/$sub"
  cp "$SRC/src/$sub"/*.py "$DST/src/$sub/"
done
[ -f "$SRC/src/__init__.py" ] && cp "$SRC/src/__init__.py" "$DST/src/" || true
cp -r "$SRC/configs/train" "$DST/configs/train"
( cd "$SRC" && find src configs/train -name '*.py' -o -name '*.json' | sort | xargs sha256sum ) > "$DST/MANIFEST.sha256"
echo "make_ref: copied $(find "$DST" -name '*.py' | wc -l) python files to $DST"

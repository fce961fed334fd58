#!/usr/bin/env bash
# Prepare a B200 box for mdi_llm_b200 (counterpart of the reference's scripts/jetson-setup.sh, which
# raised clocks / swap on a Jetson TX2).  Nothing here changes GPU clocks.  Steps: check the
# toolchain, build the sm_100a kernel library in-tree, check peer access between the GPUs, run the
# CPU test-suite.  See docs/setup-b200.md.
set -euo pipefail
cd "$(dirname "$0")/.."

echo "== toolchain"
command -v nvcc >/dev/null || { echo "nvcc not found (CUDA >= 12.8 needed for sm_100a)"; exit 1; }
nvcc --version | tail -2
python - <<'PY'
import torch
print("torch", torch.__version__, "cuda", torch.version.cuda, "gpus", torch.cuda.device_count())
PY

echo "== kernels (sm_100a)"
python -m mdi_llm_b200.ops.build

if python -c 'import torch,sys; sys.exit(0 if torch.cuda.device_count() > 1 else 1)'; then
  echo "== NVLink peer access"
  python - <<'PY'
import torch
n = torch.cuda.device_count()
bad = [(i, j) for i in range(n) for j in range(n) if i != j and not torch.cuda.can_device_access_peer(i, j)]
print("peer access: all pairs ok" if not bad else f"NO peer access for {bad}: the fused hop needs it (NCCL hop still works)")
PY
  nvidia-smi topo -m || true
fi

echo "== tests (CPU)"
python -m pytest tests -x -q -m "not gpu"
if python -c 'import torch,sys; sys.exit(0 if torch.cuda.is_available() else 1)'; then
  echo "== tests (GPU)"
  python -m pytest tests -x -q -m gpu
fi

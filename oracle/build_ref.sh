#!/usr/bin/env bash
# Compile the REFERENCE's own CUDA kernels, unmodified, from the sources where they lie under
# /root/reference, into oracle/_ref/ (git-ignored; travels to the GPU box with gpurun).
# TEST INFRASTRUCTURE: the GPU-side comparator for index-exact parity (tests/test_ref_kernels_gpu.py)
# and for the "vs reference pointops" microbench.  Only the *.cu kernel files are built (their
# `extern "C" *_launcher` entry points take raw device pointers); the at::Tensor shims (*.cpp, which
# include the long-removed THC/THC.h) are not needed.  Torch headers are on the include path only
# because the reference's *_kernel.h files include them; nothing from torch is linked.
set -euo pipefail
REF=${REPSURF_REFERENCE:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
[ -d "$REF/classification/modules/pointops/src" ] || { echo "reference not present at $REF: keeping prebuilt files"; exit 0; }
mkdir -p "$OUT"
TORCH_INC=$(python - <<'PY'
import torch.utils.cpp_extension as c
print(" ".join("-I" + p for p in c.include_paths()))
PY
)
PY_INC=$(python -c "import sysconfig; print('-I' + sysconfig.get_paths()['include'])")
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O2 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared -std=c++17 -ccbin /usr/bin/g++ $TORCH_INC $PY_INC -D_GLIBCXX_USE_CXX11_ABI=1 -w"
build() {  # tree name
  local src="$REF/$1/modules/pointops/src"
  local files=$(ls "$src"/*/*_cuda_kernel.cu)
  $NVCC $FLAGS -o "$OUT/libref_pointops_$2.so" $files
  echo "built $OUT/libref_pointops_$2.so"
}
build classification cls &
build segmentation seg &
wait

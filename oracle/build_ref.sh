#!/usr/bin/env bash
# Build recipe for oracle/_ref: the UNMODIFIED reference CUDA extension (xinhao-luo/ClusterFusion), compiled from the
# sources where they lie under /root/reference for sm_100a.  Test/measurement infrastructure only:
#   - tests/test_gpu_ref_kernel.py compares our kernels with the reference's own kernels on the same inputs (B200),
#   - bench.py reports its per-layer time beside ours ("reference_gpu_kernel": the recompiled sm_90 kernel = the
#     baseline this repo is here to beat).
# Nothing under clusterfusion_b200/ or clusterfusion/ imports it.  The reference's setup.py refuses SM 10.0
# (/root/reference/setup.py:5-15) and is not run; this script passes the same source list and macro as its sm90a
# branch (setup.py:27-36, :51-60) straight to nvcc with -gencode arch=compute_100a,code=sm_100a.
# Outputs only into oracle/_ref/ (git-ignored, travels to the GPU box).  No reference source is copied.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
PY=${PYTHON:-python}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
[ -d "$REF/include/H100/llama" ] || { echo "reference sources not found under $REF (GPU box?): keeping prebuilt $OUT"; exit 0; }
mkdir -p "$OUT/obj"
MOD=_clusterfusion_ref
SO="$OUT/$MOD$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
SRCS=(include/pybind.cpp
      include/H100/llama/llama_kernel_dispatch.cu
      include/H100/llama/llama_kernel_sglang_dispatch.cu
      include/H100/llama/llama_kernel_batch_sglang_dispatch.cu
      include/H100/deepseek/deepseek_kernel_dispatch.cu
      include/H100/norm/norm_kernel_dispatch.cu)
if [ -f "$SO" ] && [ -z "${FORCE:-}" ]; then
  newest=$(ls -t "$SO" "${SRCS[@]/#/$REF/}" | head -1)
  [ "$newest" = "$SO" ] && { echo "oracle/_ref up to date: $SO"; exit 0; }
fi
read -r TORCH_INC TORCH_LIB PY_INC ABI < <($PY - <<'PYEOF'
import os, sysconfig, torch
from torch.utils import cpp_extension as ce
inc = " ".join("-I" + p for p in ce.include_paths("cuda"))
print(inc.replace(" ", "@"), os.path.join(os.path.dirname(torch.__file__), "lib"), sysconfig.get_paths()["include"],
      int(torch._C._GLIBCXX_USE_CXX11_ABI))
PYEOF
)
TORCH_INC=${TORCH_INC//@/ }
COMMON="-O3 -std=c++17 -DCOMPILE_SM90 -DTORCH_EXTENSION_NAME=$MOD -DTORCH_API_INCLUDE_EXTENSION_H -D_GLIBCXX_USE_CXX11_ABI=$ABI $TORCH_INC -I$PY_INC"
pids=()
objs=()
for s in "${SRCS[@]}"; do
  o="$OUT/obj/$(basename "${s%.*}").o"
  objs+=("$o")
  if [[ $s == *.cu ]]; then
    $NVCC $COMMON -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC -c "$REF/$s" -o "$o" &
  else
    g++ $COMMON -fPIC -c "$REF/$s" -o "$o" &
  fi
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -o "$SO" "${objs[@]}" -L"$TORCH_LIB" -L/usr/local/cuda/lib64 -L/usr/local/cuda/lib64/stubs \
    -Wl,-rpath,"$TORCH_LIB" -Wl,-rpath,/usr/local/cuda/lib64 \
    -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python -lcudart -lcuda
rm -rf "$OUT/obj"
echo "built $SO"

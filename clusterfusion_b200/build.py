"""In-tree build of the two native artefacts:

  clusterfusion_b200/libclusterfusion_b200.so   torch-free C ABI + sm_100a kernels   (nvcc, ~1 min)
  clusterfusion/_clusterfusion*.so              PyTorch C++ extension over the C ABI  (g++, ~1 min: torch headers)

The extension lands where the reference's setup.py puts its own (`clusterfusion._clusterfusion`,
/root/reference/setup.py:48) so `from clusterfusion import llama_decoder_layer` resolves the same way; it finds the
kernel library through an $ORIGIN-relative rpath, in the source tree and in an installed copy alike (setup.py at the
repo root installs both packages side by side).

The reference builds one CUDAExtension whose every .cu includes torch/extension.h
(/root/reference/setup.py:40-63, ~3.5 min per translation unit) and refuses any GPU but SM 9.0 /
12.0 (setup.py:5-15).  Here the kernel TU is torch-free and the target is fixed: sm_100a.
"""
from __future__ import annotations

import os
import shlex
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
ROOT = HERE.parent
LIB = HERE / "libclusterfusion_b200.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
GENCODE = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _ext_path() -> Path:
    return ROOT / "clusterfusion" / ("_clusterfusion" + sysconfig.get_config_var("EXT_SUFFIX"))


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print("+", " ".join(shlex.quote(str(c)) for c in cmd), flush=True)
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError(f"build step failed: {cmd[0]}")
    return r.stdout + r.stderr


def build_cabi(force: bool = False, verbose: bool = True) -> Path:
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + sorted((ROOT / "include").glob("*"))
    if not force and _newer(LIB, srcs):
        return LIB
    _run([NVCC, *GENCODE, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-shared",
          "-o", LIB, CSRC / "llama_decoder.cu", "-lcudart"], verbose)
    return LIB


def build_ext(force: bool = False, verbose: bool = True) -> Path:
    import torch
    from torch.utils import cpp_extension as ce
    out = _ext_path()
    srcs = [CSRC / "pybind.cpp", ROOT / "include" / "clusterfusion_b200.h"]
    if not force and _newer(out, srcs) and LIB.exists() and out.stat().st_mtime >= 0:
        return out
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    libdirs = ce.library_paths("cuda")
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DTORCH_EXTENSION_NAME=_clusterfusion",
           "-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           *inc, CSRC / "pybind.cpp", "-o", out,
           f"-L{HERE}", "-lclusterfusion_b200", "-Wl,-rpath,$ORIGIN/../clusterfusion_b200",
           *[f"-L{d}" for d in libdirs], *[f"-Wl,-rpath,{d}" for d in libdirs],
           "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python", "-lcudart"]
    _run(cmd, verbose)
    return out


def build_all(force: bool = False, verbose: bool = True):
    return build_cabi(force, verbose), build_ext(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print("built:", LIB.name, _ext_path().name)

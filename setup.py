"""pip-installable build of the drop-in package: `pip install --no-build-isolation .`

Builds, with the same in-tree recipe `__graft_entry__.build()` uses (clusterfusion_b200/build.py),
  clusterfusion_b200/libclusterfusion_b200.so   nvcc -gencode arch=compute_100a,code=sm_100a  (torch-free C ABI + kernels)
  clusterfusion/_clusterfusion*.so              g++ against the installed torch               (the reference's module name)
and installs the two packages side by side, so `from clusterfusion import llama_decoder_layer` resolves to the native
module `clusterfusion._clusterfusion` exactly as with the reference's own setup.py (/root/reference/setup.py:40-63), which
refuses any GPU but SM 9.0 / 12.0 (:5-15); this one targets sm_100a only.  The extension finds the kernel library through
the rpath $ORIGIN/../clusterfusion_b200, in the source tree and in site-packages alike.
"""
import importlib.util
from pathlib import Path

from setuptools import setup
from setuptools.command.build_py import build_py
from setuptools.dist import Distribution

ROOT = Path(__file__).resolve().parent


class BuildNativeThenPy(build_py):
    def run(self):
        spec = importlib.util.spec_from_file_location("_cfb200_build", ROOT / "clusterfusion_b200" / "build.py")
        b = importlib.util.module_from_spec(spec)      # loaded by path: the package refuses to import before the build
        spec.loader.exec_module(b)
        b.build_all(force=False, verbose=True)
        super().run()


class BinaryDistribution(Distribution):
    def has_ext_modules(self):          # platform wheel: the package carries prebuilt shared objects
        return True


setup(
    name="clusterfusion",
    version="0.3.0",
    description="B200-native (sm_100a) fused Llama decoder-layer operators behind the ClusterFusion operator names",
    packages=["clusterfusion", "clusterfusion_b200"],
    package_data={"clusterfusion": ["_clusterfusion*.so"],
                  "clusterfusion_b200": ["libclusterfusion_b200.so", "csrc/*.cu", "csrc/*.cuh", "csrc/*.cpp"]},
    include_package_data=False,
    python_requires=">=3.10",
    install_requires=["torch"],
    cmdclass={"build_py": BuildNativeThenPy},
    distclass=BinaryDistribution,
    zip_safe=False,
)

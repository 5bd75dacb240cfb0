"""Packaging (/root/reference/setup.py:40-63): the native module must be importable under the
reference's name `clusterfusion._clusterfusion`, and `pip install .` must produce a self-contained copy that imports
from outside the source tree."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_native_module_has_the_reference_name():
    import clusterfusion
    import clusterfusion_b200
    assert clusterfusion._clusterfusion.__name__ == "clusterfusion._clusterfusion"
    for name in ("llama_decoder_layer", "llama_decoder_layer_sglang", "llama_decoder_layer_batch_decode_sglang",
                 "deepseek_decoder_layer", "rmsnorm"):
        assert getattr(clusterfusion, name) is getattr(clusterfusion._clusterfusion, name) is getattr(clusterfusion_b200, name)


def test_pip_install_into_a_scratch_prefix(tmp_path):
    target = tmp_path / "site"
    r = subprocess.run([sys.executable, "-m", "pip", "install", "--no-build-isolation", "--no-index", "--no-deps", "--quiet",
                        "--target", str(target), str(ROOT)], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout + r.stderr
    code = ("import clusterfusion, clusterfusion_b200; from clusterfusion_b200 import cabi; "
            "assert clusterfusion.abi_version() == cabi.load().cf_abi_version(); "
            "print(clusterfusion._clusterfusion.__file__); print(cabi.LIB_PATH)")
    env = dict(os.environ, PYTHONPATH=str(target))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    ext, lib = r.stdout.strip().splitlines()[-2:]
    assert ext.startswith(str(target)) and lib.startswith(str(target)), (ext, lib)

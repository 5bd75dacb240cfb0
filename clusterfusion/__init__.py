"""Drop-in package for the reference's `clusterfusion` (/root/reference/clusterfusion/__init__.py:6-16): the native module
is `clusterfusion._clusterfusion` -- the extension name the reference's setup.py builds (setup.py:48) -- and every public
name of it is re-exported here, so `from clusterfusion import llama_decoder_layer` (chat/llama/model.py:19,
tests/test_llama.py:6) resolves to the B200-native operators.  The kernels live in
`clusterfusion_b200/libclusterfusion_b200.so` (torch-free C ABI); the extension is a thin shim over it.
There is no Python or CPU fallback: without the built extension this import raises ImportError, like the reference's."""
import atexit as _atexit

try:
    from . import _clusterfusion
except ImportError as e:  # same behaviour as /root/reference/clusterfusion/__init__.py:6-12
    raise ImportError(
        "Failed to import the clusterfusion native extension. Build it in-tree with "
        "`python clusterfusion_b200/build.py` (needs nvcc for sm_100a) or `pip install .`; there is no fallback path."
    ) from e

for _attr in dir(_clusterfusion):
    if not _attr.startswith("_"):
        globals()[_attr] = getattr(_clusterfusion, _attr)

__all__ = [a for a in dir(_clusterfusion) if not a.startswith("_")]
# cached CUDA workspaces must be released while the CUDA context still exists, not by static destructors at process exit
_atexit.register(_clusterfusion._release_workspaces)
del _attr, _atexit

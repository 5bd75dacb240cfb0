"""Drop-in name for the reference package: ``from clusterfusion import llama_decoder_layer`` (the import
the reference's chat/llama/model.py:19 and tests/test_llama.py:6 perform) resolves to the B200-native
operators in ``clusterfusion_b200``.  Raises ImportError if the native extension is not built."""
from clusterfusion_b200 import *  # noqa: F401,F403
from clusterfusion_b200 import __all__  # noqa: F401

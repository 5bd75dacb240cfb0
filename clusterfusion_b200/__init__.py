"""clusterfusion_b200 -- B200-native (sm_100a) fused Llama decoder attention half-layer.

Public operator surface = the reference's (``/root/reference/clusterfusion/__init__.py:6-16`` re-exports
every public name of its native module; ``include/pybind.cpp:108-123`` defines them):

    llama_decoder_layer(input, weight_qkv, weight_o, k_cache, v_cache, rms_w, cos, sin) -> (o, k, v)
    llama_decoder_layer_sglang(input, residual, weight_qkv, weight_o, k_cache, v_cache, rms_w, eps,
                               cos, sin) -> (o, residual, k, v)
    llama_decoder_layer_batch_decode_sglang(output, residual_output, input, residual, weight_qkv,
                               weight_o, paged_kv_indptr, paged_kv_indices, k_cache_ptrs,
                               v_cache_ptrs, layer_id, rms_w, eps, positions, cos_sin) -> None

and, from the reference's sm90 module (``include/pybind.cpp:45-64``, ``:113-114``):

    deepseek_decoder_layer(input, weight_q_nope, weight_q_pe, weight_uk, weight_kv_nope, weight_k_pe,
                           weight_uv, weight_o, ckv_cache, rms_input_weight, rms_ckv_weight, cos, sin) -> o
    rmsnorm(input, weight) -> out

plus operators the reference does not have (``llama_ffn_layer``, ``deepseek_decoder_layer_ex``, ``set_pdl``).

Everything is native: a torch-free C-ABI library (``libclusterfusion_b200.so``, the CUDA kernels) and a
thin PyTorch C++ extension (``clusterfusion._clusterfusion``, the reference's module name).  There is no Python or CPU fallback -- importing
this package without the built extension raises ImportError, exactly like the reference package.
"""
import importlib as _importlib

try:
    _ext = _importlib.import_module("clusterfusion._clusterfusion")
except ImportError as e:  # same behaviour as /root/reference/clusterfusion/__init__.py:6-12
    raise ImportError(
        "Failed to import the clusterfusion native extension. Build it in-tree with "
        "`python clusterfusion_b200/build.py` (needs nvcc for sm_100a) or `pip install .`; there is no fallback path."
    ) from e

for _attr in dir(_ext):
    if not _attr.startswith("_"):
        globals()[_attr] = getattr(_ext, _attr)

__all__ = [a for a in dir(_ext) if not a.startswith("_")]
del _importlib, _attr

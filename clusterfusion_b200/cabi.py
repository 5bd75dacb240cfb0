"""ctypes binding of the C ABI in include/clusterfusion_b200.h.

This is the thinnest possible host layer over ``libclusterfusion_b200.so``: raw device
pointers in, one kernel launch out.  The pybind extension (csrc/pybind.cpp) sits on the
same entry points; this module exists so tests and bench.py can drive the C ABI directly
(``data_ptr()`` integers, no torch types crossing the boundary).

There is no CPU fallback: if the shared library is missing the import raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
# CF_LIB_PATH: explicit override for kernel experiments (tools/); the product always loads the in-tree library
LIB_PATH = Path(os.environ["CF_LIB_PATH"]) if os.environ.get("CF_LIB_PATH") else _HERE / "libclusterfusion_b200.so"

CF_VARIANT_CHAT, CF_VARIANT_SGLANG, CF_VARIANT_PAGED = 0, 1, 2
CF_FLAG_OUT_FP32_PARTIAL = 0x1
CF_FLAG_PDL = 0x2
CF_FLAG_LL_OUT = 0x8
CF_FLAG_PER_REQUEST = 0x10
CF_DS_FLAG_ROPE_SCORES = 0x100

EXPORTED_SYMBOLS = (
    "cf_abi_version",
    "cf_sizeof_llama_args",
    "cf_sizeof_ffn_args",
    "cf_last_error_string",
    "cf_llama_workspace_bytes",
    "cf_rmsnorm_launch",
    "cf_tp_exchange_bytes",
    "cf_ipc_alloc",
    "cf_ipc_open",
    "cf_ipc_close",
    "cf_ipc_free",
    "cf_llama_algorithmic_bytes",
    "cf_llama_decoder_layer_launch",
    "cf_llama_ffn_launch",
    "cf_deepseek_workspace_bytes",
    "cf_deepseek_decoder_layer_launch",
    "cf_sizeof_deepseek_args",
    "cf_test_cluster_reduce",
    "cf_workspace_status",
    "cf_workspace_clear_status",
    "cf_debug_tensor_map_encodes",
)


class CfLlamaArgs(C.Structure):
    _fields_ = [
        ("variant", C.c_int32),
        ("flags", C.c_uint32),
        ("hidden", C.c_int32),
        ("n_q_heads", C.c_int32),
        ("n_kv_heads", C.c_int32),
        ("head_dim", C.c_int32),
        ("batch", C.c_int32),
        ("kv_len", C.c_uint32),
        ("layer_id", C.c_int32),
        ("eps", C.c_float),
        ("x", C.c_void_p),
        ("residual_in", C.c_void_p),
        ("w_qkv", C.c_void_p),
        ("w_o", C.c_void_p),
        ("rms_w", C.c_void_p),
        ("out", C.c_void_p),
        ("residual_out", C.c_void_p),
        ("k_new", C.c_void_p),
        ("v_new", C.c_void_p),
        ("k_cache", C.c_void_p),
        ("v_cache", C.c_void_p),
        ("indptr", C.c_void_p),
        ("indices", C.c_void_p),
        ("k_pool_ptrs", C.c_void_p),
        ("v_pool_ptrs", C.c_void_p),
        ("positions", C.c_void_p),
        ("cos", C.c_void_p),
        ("sin", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_batch", C.c_int32),
        ("tp_rank", C.c_int32),
        ("tp_world", C.c_int32),
        ("tp_peer", C.c_void_p * 8),
    ]


class CfFfnArgs(C.Structure):
    _fields_ = [
        ("flags", C.c_uint32),
        ("hidden", C.c_int32),
        ("ffn", C.c_int32),
        ("eps", C.c_float),
        ("x", C.c_void_p),
        ("residual_in", C.c_void_p),
        ("w_gate_up", C.c_void_p),
        ("w_down_t", C.c_void_p),
        ("rms_w", C.c_void_p),
        ("out", C.c_void_p),
        ("residual_out", C.c_void_p),
        ("workspace", C.c_void_p),
        ("workspace_batch", C.c_int32),
    ]


class CfDeepseekArgs(C.Structure):
    _fields_ = [
        ("flags", C.c_uint32),
        ("hidden", C.c_int32),
        ("n_heads", C.c_int32),
        ("seq_len", C.c_int32),
        ("eps", C.c_float),
        ("x", C.c_void_p),
        ("w_q_nope", C.c_void_p),
        ("w_q_pe", C.c_void_p),
        ("w_uk", C.c_void_p),
        ("w_kv_nope", C.c_void_p),
        ("w_k_pe", C.c_void_p),
        ("w_uv", C.c_void_p),
        ("w_o", C.c_void_p),
        ("ckv_cache", C.c_void_p),
        ("rms_input_w", C.c_void_p),
        ("rms_ckv_w", C.c_void_p),
        ("cos", C.c_void_p),
        ("sin", C.c_void_p),
        ("out", C.c_void_p),
        ("ckv_new", C.c_void_p),
        ("k_pe_new", C.c_void_p),
        ("workspace", C.c_void_p),
    ]


class CfError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"clusterfusion_b200 C ABI error {code}: {msg}")
        self.code = code


_lib = None


def load() -> C.CDLL:
    """dlopen the C-ABI library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc -gencode arch=compute_100a,code=sm_100a).  There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else C.DEFAULT_MODE)
    lib.cf_abi_version.restype = C.c_int
    lib.cf_sizeof_llama_args.restype = C.c_size_t
    lib.cf_sizeof_ffn_args.restype = C.c_size_t
    # CF_LIB_PATH experiments (tools/) may load an older build whose FFN / DeepSeek structs differ; the product never sets it
    experiment = bool(os.environ.get("CF_LIB_PATH")) and os.environ.get("CF_SKIP_ABI_CHECK") == "1"
    if not experiment and (lib.cf_sizeof_llama_args() != C.sizeof(CfLlamaArgs) or lib.cf_sizeof_ffn_args() != C.sizeof(CfFfnArgs)):
        raise ImportError(f"clusterfusion_b200.cabi: struct mirror out of date (CfLlamaArgs {C.sizeof(CfLlamaArgs)} vs "
                          f"{lib.cf_sizeof_llama_args()}, CfFfnArgs {C.sizeof(CfFfnArgs)} vs {lib.cf_sizeof_ffn_args()}): rebuild")
    lib.cf_sizeof_deepseek_args.restype = C.c_size_t
    if lib.cf_sizeof_deepseek_args() != C.sizeof(CfDeepseekArgs):
        raise ImportError(f"clusterfusion_b200.cabi: CfDeepseekArgs mirror out of date ({C.sizeof(CfDeepseekArgs)} vs "
                          f"{lib.cf_sizeof_deepseek_args()}): rebuild")
    lib.cf_deepseek_workspace_bytes.restype = C.c_size_t
    lib.cf_deepseek_decoder_layer_launch.restype = C.c_int
    lib.cf_deepseek_decoder_layer_launch.argtypes = [C.POINTER(CfDeepseekArgs), C.c_void_p]
    lib.cf_last_error_string.restype = C.c_char_p
    lib.cf_llama_workspace_bytes.restype = C.c_size_t
    lib.cf_llama_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.cf_llama_algorithmic_bytes.restype = C.c_uint64
    lib.cf_llama_algorithmic_bytes.argtypes = [C.POINTER(CfLlamaArgs), C.c_uint64]
    lib.cf_llama_decoder_layer_launch.restype = C.c_int
    lib.cf_llama_decoder_layer_launch.argtypes = [C.POINTER(CfLlamaArgs), C.c_void_p]
    lib.cf_llama_ffn_launch.restype = C.c_int
    lib.cf_llama_ffn_launch.argtypes = [C.POINTER(CfFfnArgs), C.c_void_p]
    lib.cf_tp_exchange_bytes.restype = C.c_size_t
    lib.cf_tp_exchange_bytes.argtypes = [C.c_int32, C.c_int32]
    lib.cf_ipc_alloc.restype = C.c_int
    lib.cf_ipc_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    lib.cf_ipc_open.restype = C.c_int
    lib.cf_ipc_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.cf_ipc_close.restype = C.c_int
    lib.cf_ipc_close.argtypes = [C.c_void_p]
    lib.cf_ipc_free.restype = C.c_int
    lib.cf_ipc_free.argtypes = [C.c_void_p]
    lib.cf_rmsnorm_launch.restype = C.c_int
    lib.cf_rmsnorm_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_uint32, C.c_void_p]
    lib.cf_test_cluster_reduce.restype = C.c_int
    lib.cf_test_cluster_reduce.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32, C.c_void_p]
    if not experiment:
        lib.cf_workspace_status.restype = C.c_int
        lib.cf_workspace_status.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        lib.cf_workspace_clear_status.restype = C.c_int
        lib.cf_workspace_clear_status.argtypes = [C.c_void_p, C.c_void_p]
        lib.cf_debug_tensor_map_encodes.restype = C.c_uint64
    _lib = lib
    return lib


def last_error() -> str:
    return load().cf_last_error_string().decode()


def check(rc: int) -> None:
    if rc != 0:
        raise CfError(rc, last_error())


def workspace_bytes(hidden: int, batch: int = 1) -> int:
    return int(load().cf_llama_workspace_bytes(hidden, batch))


def algorithmic_bytes(args: CfLlamaArgs, total_kv_rows: int) -> int:
    return int(load().cf_llama_algorithmic_bytes(C.byref(args), total_kv_rows))


def workspace_status(workspace_ptr: int, stream: int = 0) -> int:
    """Sticky error word of a workspace (0 = ok; non-zero = an in-kernel exchange poll timed out since the last clear).
    Synchronises `stream`."""
    st = C.c_uint32(0)
    check(load().cf_workspace_status(C.c_void_p(workspace_ptr), C.c_void_p(stream), C.byref(st)))
    return int(st.value)


def workspace_clear_status(workspace_ptr: int, stream: int = 0) -> None:
    check(load().cf_workspace_clear_status(C.c_void_p(workspace_ptr), C.c_void_p(stream)))


def tensor_map_encodes() -> int:
    """cuTensorMapEncodeTiled calls made by the library so far in this process."""
    return int(load().cf_debug_tensor_map_encodes())


def launch(args: CfLlamaArgs, stream: int = 0) -> None:
    """One fused-kernel launch on CUDA stream handle `stream` (0 = legacy default stream)."""
    check(load().cf_llama_decoder_layer_launch(C.byref(args), C.c_void_p(stream)))


def launch_deepseek(args: CfDeepseekArgs, stream: int = 0) -> None:
    """cf_deepseek_decoder_layer_launch; raises CfError on a non-zero return code."""
    check(load().cf_deepseek_decoder_layer_launch(C.byref(args), C.c_void_p(stream)))


def launch_ffn(args: CfFfnArgs, stream: int = 0) -> None:
    """One fused FFN half-layer launch on CUDA stream handle `stream`."""
    check(load().cf_llama_ffn_launch(C.byref(args), C.c_void_p(stream)))


def test_cluster_reduce(in_ptr: int, out_ptr: int, n: int, cluster_size: int, n_clusters: int,
                        stage: int, repeats: int, stream: int = 0) -> None:
    check(load().cf_test_cluster_reduce(C.c_void_p(in_ptr), C.c_void_p(out_ptr), n, cluster_size,
                                        n_clusters, stage, repeats, C.c_void_p(stream)))

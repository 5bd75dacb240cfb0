"""Test helper: marshal torch CUDA tensors into a CfLlamaArgs and launch through the raw C ABI (ctypes).
Nothing here computes anything -- it only passes pointers."""
import torch

from clusterfusion_b200 import cabi

_ws = {}


def workspace(hidden, batch, device):
    key = (hidden, batch, str(device))
    if key not in _ws:
        _ws[key] = torch.zeros(cabi.workspace_bytes(hidden, batch), dtype=torch.uint8, device=device)
    return _ws[key]


def _p(t):
    return None if t is None else t.data_ptr()


def stream_handle():
    return torch.cuda.current_stream().cuda_stream


def chat(x, w_qkv, w_o, k_cache, v_cache, rms_w, cos, sin, eps=1e-6, flags=0):
    hidden = x.shape[-1]
    H = hidden // 128
    o = torch.empty(1, hidden, dtype=torch.float16, device=x.device)
    k = torch.empty(1, H, 128, dtype=torch.float16, device=x.device)
    v = torch.empty(1, H, 128, dtype=torch.float16, device=x.device)
    a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_CHAT, flags=flags, hidden=hidden, n_q_heads=H, n_kv_heads=H, head_dim=128,
                         batch=1, kv_len=k_cache.shape[0], eps=eps, x=_p(x), w_qkv=_p(w_qkv), w_o=_p(w_o),
                         rms_w=_p(rms_w), out=_p(o), k_new=_p(k), v_new=_p(v), k_cache=_p(k_cache),
                         v_cache=_p(v_cache), cos=_p(cos), sin=_p(sin),
                         workspace=_p(workspace(hidden, 1, x.device)))
    cabi.launch(a, stream_handle())
    return o, k, v


def sglang(x, residual, w_qkv, w_o, k_cache, v_cache, rms_w, eps, cos, sin, n_heads, n_kv_heads=None,
           residual_out=None, flags=0):
    n_kv_heads = n_kv_heads or n_heads
    hidden = x.shape[-1]
    o = torch.empty(1, hidden, dtype=torch.float16, device=x.device)
    k = torch.empty(1, n_kv_heads, 128, dtype=torch.float16, device=x.device)
    v = torch.empty(1, n_kv_heads, 128, dtype=torch.float16, device=x.device)
    if residual_out is None:
        residual_out = torch.empty_like(residual)
    a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_SGLANG, flags=flags, hidden=hidden, n_q_heads=n_heads, n_kv_heads=n_kv_heads,
                         head_dim=128, batch=1, kv_len=k_cache.shape[0], eps=eps, x=_p(x), residual_in=_p(residual),
                         residual_out=_p(residual_out), w_qkv=_p(w_qkv), w_o=_p(w_o), rms_w=_p(rms_w), out=_p(o),
                         k_new=_p(k), v_new=_p(v), k_cache=_p(k_cache), v_cache=_p(v_cache), cos=_p(cos), sin=_p(sin),
                         workspace=_p(workspace(hidden, 1, x.device)))
    cabi.launch(a, stream_handle())
    return o, residual_out, k, v


def paged(out, residual_out, x, residual, w_qkv, w_o, indptr, indices, k_ptrs, v_ptrs, layer_id, rms_w, eps,
          positions, cos_sin, n_heads, n_kv_heads=None, flags=0):
    n_kv_heads = n_kv_heads or n_heads
    bs, hidden = x.shape
    a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_PAGED, flags=flags, hidden=hidden, n_q_heads=n_heads, n_kv_heads=n_kv_heads,
                         head_dim=128, batch=bs, layer_id=layer_id, eps=eps, x=_p(x), residual_in=_p(residual),
                         residual_out=_p(residual_out), w_qkv=_p(w_qkv), w_o=_p(w_o), rms_w=_p(rms_w), out=_p(out),
                         indptr=_p(indptr), indices=_p(indices), k_pool_ptrs=_p(k_ptrs), v_pool_ptrs=_p(v_ptrs),
                         positions=_p(positions), cos=_p(cos_sin), workspace=_p(workspace(hidden, bs, x.device)))
    cabi.launch(a, stream_handle())

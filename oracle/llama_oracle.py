"""CPU oracle for the fused Llama decoder attention half-layer.

TEST INFRASTRUCTURE ONLY.  Nothing under ``clusterfusion_b200/`` (the product)
may import this module.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it, and only as
the checker / the timed CPU baseline -- never as the thing shipped.

It restates, in plain CPU PyTorch, what the reference computes on the hot path

    RMSNorm -> QKV projection -> RoPE -> decode attention over
    (KV cache ++ current token) -> O projection  (-> residual bookkeeping)

for the three public signatures of the reference:

* ``chat_layer``    -- 8-arg ``llama_decoder_layer``
                       (/root/reference/include/pybind.cpp:3-12,
                        include/H100/llama/kernel.cuh:20-619; eager twin
                        chat/llama/model.py:54-79 RMSNorm, :134-163 RoPE,
                        :376-405 attention, :292-328 fused-weight layout)
* ``sglang_layer``  -- 10-arg ``llama_decoder_layer_sglang``
                       (include/pybind.cpp:14-25, kernel_sglang.cuh:20-633;
                        pure-torch twin tests/test_llama_tilelang.py:18-49)
* ``paged_layer``   -- 15-arg ``llama_decoder_layer_batch_decode_sglang``
                       (include/pybind.cpp:27-43,
                        kernel_batch_sglang.cuh:43-664; paging semantics
                        :118-123, :322-344)

Pinning (see tests/test_oracle_golden.py, oracle/gen_golden.py): the oracle is
checked against fixtures produced in the build container by *executing the
reference's own Python* -- ``tests/test_llama_tilelang.py::reference`` as is,
and ``chat/llama/model.py``'s ``RMSNorm`` / ``precompute_freqs_cis`` /
``apply_rotary_emb`` / ``Attention`` (eager branch) with the two absent
third-party packages stubbed (fairscale layers -> ``nn.Linear``;
``flashinfer.single_decode_with_kv_cache`` -> softmax(QK^T/sqrt(d))V, its
published definition; both packages are unpinned by the reference).

Two arithmetic flavours:

* ``mode="fp32"``  -- fp16 inputs, every intermediate in fp32, outputs rounded
  once.  This is what tests/test_llama_tilelang.py::reference does.
* ``mode="eager"`` -- fp16 rounding at every point where the eager fp16 model
  materialises an fp16 tensor (normed x twice, q/k/v, RoPE outputs, attention
  output, O projection).  This is the "reference eager-PyTorch layer" of the
  north-star; accumulation inside each op is fp32, as cuBLAS / flashinfer do.

The CUDA kernel is required to match either flavour to rtol = atol = 1e-3.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

HEAD_DIM = 128


@dataclass(frozen=True)
class LayerShape:
    hidden: int = 4096
    n_heads: int = 32
    n_kv_heads: int = 32
    head_dim: int = HEAD_DIM

    @property
    def q_dim(self) -> int:
        return self.n_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.n_kv_heads * self.head_dim

    @property
    def qkv_dim(self) -> int:
        return self.q_dim + 2 * self.kv_dim


LLAMA2_7B = LayerShape(4096, 32, 32)
LLAMA3_8B = LayerShape(4096, 32, 8)
LLAMA2_70B = LayerShape(8192, 64, 8)


def _r16(t: torch.Tensor, mode: str) -> torch.Tensor:
    """Round to fp16 and come back to fp32 in eager mode; identity in fp32 mode."""
    return t.half().float() if mode == "eager" else t


# --------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------
def rmsnorm(h: torch.Tensor, w: torch.Tensor, eps: float, mode: str) -> torch.Tensor:
    """chat/llama/model.py:54-79: ``_norm(x.float()).type_as(x) * weight``.

    ``h`` fp32 [..., hidden] (already holding fp16-representable values in eager
    mode), ``w`` fp16/fp32 [hidden].  Returns fp32.
    """
    n = h * torch.rsqrt(h.pow(2).mean(-1, keepdim=True) + eps)
    n = _r16(n, mode)
    return _r16(n * w.float(), mode)


def rope_gptj(t: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """Interleaved-pair rotation, chat/llama/model.py:134-163 (complex multiply).

    ``t`` [heads, D]; ``cos``/``sin`` [D] pair-repeated (model.py:278-280).
    """
    te, to = t[..., 0::2], t[..., 1::2]
    c, s = cos[..., 0::2], sin[..., 0::2]
    out = torch.empty_like(t)
    out[..., 0::2] = te * c - to * s
    out[..., 1::2] = to * c + te * s
    return out


def rope_neox(t: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """Rotate-half, tests/test_llama_tilelang.py:31-34; ``cos``/``sin`` [D/2]."""
    half = t.shape[-1] // 2
    t1, t2 = t[..., :half], t[..., half:]
    return torch.cat([t1 * cos - t2 * sin, t2 * cos + t1 * sin], dim=-1)


def decode_attention(q: torch.Tensor, K: torch.Tensor, V: torch.Tensor) -> torch.Tensor:
    """softmax(q K^T / sqrt(D)) V for one query token.

    q [Hq, D]; K, V [S, Hkv, D] (S >= 1, current token included).  GQA: Q head h
    uses KV head h // (Hq // Hkv) (chat/llama/model.py:166-175 ``repeat_kv``).
    fp32 throughout (flashinfer.single_decode_with_kv_cache, model.py:261-268,
    accumulates in fp32; eager prefill path :254-260 uses softmax(...float())).
    """
    hq, d = q.shape
    hkv = K.shape[1]
    rep = hq // hkv
    qg = q.view(hkv, rep, d)
    scores = torch.einsum("grd,sgd->grs", qg, K) / math.sqrt(d)
    probs = torch.softmax(scores, dim=-1)
    o = torch.einsum("grs,sgd->grd", probs, V)
    return o.reshape(hq, d)


def _core(
    h: torch.Tensor,            # fp32 [hidden] -- input to the norm
    wq: torch.Tensor,           # [Hq*D, hidden]   nn.Linear layout ([out, in])
    wk: torch.Tensor,           # [Hkv*D, hidden]
    wv: torch.Tensor,           # [Hkv*D, hidden]
    wo: torch.Tensor,           # [hidden, Hq*D]
    k_cache: torch.Tensor,      # [S, Hkv*D]
    v_cache: torch.Tensor,      # [S, Hkv*D]
    rms_w: torch.Tensor,
    eps: float,
    cos: torch.Tensor,
    sin: torch.Tensor,
    rope: str,
    shape: LayerShape,
    mode: str,
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    D = shape.head_dim
    n = rmsnorm(h, rms_w, eps, mode)
    q = _r16(wq.float() @ n, mode).view(shape.n_heads, D)
    k = _r16(wk.float() @ n, mode).view(shape.n_kv_heads, D)
    v = _r16(wv.float() @ n, mode).view(shape.n_kv_heads, D)
    cos = cos.float().reshape(-1)
    sin = sin.float().reshape(-1)
    if rope == "gptj":
        q, k = rope_gptj(q, cos[:D], sin[:D]), rope_gptj(k, cos[:D], sin[:D])
    elif rope == "neox":
        q, k = rope_neox(q, cos[: D // 2], sin[: D // 2]), rope_neox(k, cos[: D // 2], sin[: D // 2])
    else:
        raise ValueError(rope)
    q, k = _r16(q, mode), _r16(k, mode)
    S = k_cache.shape[0]
    K = torch.cat([k_cache.float().view(S, shape.n_kv_heads, D), k[None]], dim=0)
    V = torch.cat([v_cache.float().view(S, shape.n_kv_heads, D), v[None]], dim=0)
    o = _r16(decode_attention(q, K, V), mode)
    out = wo.float() @ o.reshape(-1)
    return out, k, v


# --------------------------------------------------------------------------------------
# the three public signatures
# --------------------------------------------------------------------------------------
def chat_layer(
    x: torch.Tensor,            # fp16 [1, hidden] or [1, 1, hidden]
    weight_qkv: torch.Tensor,   # fp16 [3*hidden, hidden] = [Wq^T; Wk^T; Wv^T]  (model.py:317-320)
    weight_o: torch.Tensor,     # fp16 [hidden, hidden]   = Wo^T                (model.py:322)
    k_cache: torch.Tensor,      # fp16 [S, hidden]
    v_cache: torch.Tensor,      # fp16 [S, hidden]
    rms_w: torch.Tensor,        # fp16 [hidden]
    cos: torch.Tensor,          # fp32 [1, D] pair-repeated (model.py:278-280)
    sin: torch.Tensor,
    *,
    n_heads: int = 32,
    eps: float = 1e-6,          # hard-coded in kernel.cuh:58
    mode: str = "fp32",
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """8-arg ``llama_decoder_layer``: returns (o [1,hidden], k [1,H,D], v [1,H,D]) fp16.

    No residual add: the caller does ``x + attention(...)`` (model.py:488-492).
    """
    hidden = x.shape[-1]
    shape = LayerShape(hidden, n_heads, n_heads, hidden // n_heads)
    wT = weight_qkv.view(3, hidden, hidden)
    out, k, v = _core(
        x.reshape(-1).float(), wT[0].t(), wT[1].t(), wT[2].t(), weight_o.t(),
        k_cache, v_cache, rms_w, eps, cos, sin, "gptj", shape, mode)
    return out.half().view(1, hidden), k.half().unsqueeze(0), v.half().unsqueeze(0)


def sglang_layer(
    x: torch.Tensor,            # fp16 [1, hidden]
    residual: torch.Tensor,     # fp16 [1, hidden]
    weight_qkv: torch.Tensor,   # fp16 [(Hq+2Hkv)*D, hidden]  nn.Linear layout
    weight_o: torch.Tensor,     # fp16 [hidden, Hq*D]         nn.Linear layout
    k_cache: torch.Tensor,      # fp16 [S, Hkv*D]
    v_cache: torch.Tensor,
    rms_w: torch.Tensor,
    eps: float,
    cos: torch.Tensor,          # fp32, first D/2 entries used (kernel_sglang.cuh:292-293)
    sin: torch.Tensor,
    *,
    n_heads: int = 32,
    n_kv_heads: Optional[int] = None,
    mode: str = "fp32",
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """10-arg ``llama_decoder_layer_sglang``: (o, residual_out, k, v), all fp16.

    residual_out = x + residual (flashinfer.fused_add_rmsnorm semantics,
    tests/test_llama.py:62); the norm is taken over that sum.
    """
    n_kv_heads = n_kv_heads or n_heads
    hidden = x.shape[-1]
    D = weight_o.shape[1] // n_heads
    shape = LayerShape(hidden, n_heads, n_kv_heads, D)
    h = x.reshape(-1).float() + residual.reshape(-1).float()
    res_out = h.half()
    if mode == "eager":
        h = res_out.float()          # kernel_sglang.cuh:100-105 norms the rounded sum
    wq, wk, wv = weight_qkv.split([shape.q_dim, shape.kv_dim, shape.kv_dim], dim=0)
    out, k, v = _core(h, wq, wk, wv, weight_o, k_cache, v_cache, rms_w, eps,
                      cos, sin, "neox", shape, mode)
    return (out.half().view(1, hidden), res_out.view(1, hidden),
            k.half().unsqueeze(0), v.half().unsqueeze(0))


def paged_layer(
    x: torch.Tensor,                 # fp16 [bs, hidden]
    residual: torch.Tensor,          # fp16 [bs, hidden]
    weight_qkv: torch.Tensor,
    weight_o: torch.Tensor,
    paged_kv_indptr: torch.Tensor,   # int32 [bs+1]
    paged_kv_indices: torch.Tensor,  # int32 [nnz]; last index of each request = slot of the new token
    k_pool: torch.Tensor,            # fp16 [num_slots, Hkv*D]  (pool of layer `layer_id`) -- updated in place
    v_pool: torch.Tensor,
    rms_w: torch.Tensor,
    eps: float,
    positions: torch.Tensor,         # int64 [bs]
    cos_sin: torch.Tensor,           # fp32 [max_pos, D] = [cos(D/2) | sin(D/2)]
    *,
    n_heads: int = 32,
    n_kv_heads: Optional[int] = None,
    mode: str = "fp32",
) -> Tuple[torch.Tensor, torch.Tensor]:
    """15-arg ``llama_decoder_layer_batch_decode_sglang`` with page_size = 1.

    Request b attends over pool rows ``indices[indptr[b] : indptr[b+1]-1]`` plus
    its own new token, whose post-RoPE K / raw V are written to pool slot
    ``indices[indptr[b+1]-1]`` (kernel_batch_sglang.cuh:118-123, :343-344).
    Returns (output [bs, hidden], residual_output [bs, hidden]); pools mutated.
    """
    bs, hidden = x.shape
    outs, ress = [], []
    D = weight_o.shape[1] // n_heads
    new_rows = []
    for b in range(bs):
        s, e = int(paged_kv_indptr[b]), int(paged_kv_indptr[b + 1]) - 1
        rows = paged_kv_indices[s:e].long()
        cs = cos_sin[int(positions[b])]
        o, r, k, v = sglang_layer(
            x[b:b + 1], residual[b:b + 1], weight_qkv, weight_o,
            k_pool[rows], v_pool[rows], rms_w, eps, cs[: D // 2], cs[D // 2: D],
            n_heads=n_heads, n_kv_heads=n_kv_heads, mode=mode)
        outs.append(o)
        ress.append(r)
        new_rows.append((int(paged_kv_indices[e]), k.reshape(-1), v.reshape(-1)))
    # writes happen after all reads: a request never reads another's new slot
    for slot, k, v in new_rows:
        k_pool[slot] = k
        v_pool[slot] = v
    return torch.cat(outs, 0), torch.cat(ress, 0)


# --------------------------------------------------------------------------------------
# standalone rmsnorm op
# --------------------------------------------------------------------------------------
def rmsnorm_op(x: torch.Tensor, w: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    """``clusterfusion.rmsnorm(input, weight)``: restates /root/reference/include/H100/norm/kernel.cuh:28-75 -- sum of
    squares, ``rsqrt(sum / hidden + eps)`` (eps 1e-6 hard-coded, :28) and ``x * rms_rcp * w`` all in fp32, one rounding to
    fp16 (:71).  The reference's own check compares against ``flashinfer.norm.rmsnorm`` (tests/test_norm.py:12), same
    definition.  x fp16 [batch, hidden], w fp16 [hidden] -> fp16 [batch, hidden]."""
    xf = x.float()
    r = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return (xf * r * w.float()).half()


# --------------------------------------------------------------------------------------
# FFN half-layer (the reference ships no fused FFN kernel, only the eager module)
# --------------------------------------------------------------------------------------
def ffn_layer(
    x: torch.Tensor,            # fp16 [1, hidden]
    residual: torch.Tensor,     # fp16 [1, hidden]
    w_gate_up: torch.Tensor,    # fp16 [2*ffn, hidden] = [W1; W3]  (nn.Linear layout; chat/llama/model.py:437-445)
    w_down_t: torch.Tensor,     # fp16 [ffn, hidden]   = W2^T       (W2 is [hidden, ffn], model.py:440-442)
    rms_w: torch.Tensor,        # fp16 [hidden]  (ffn_norm weight)
    eps: float,
    mode: str = "fp32",
):
    """h = x + residual; out = W2 (silu(W1 n) * (W3 n)), n = rmsnorm(h) * w  (model.py:447-448 SwiGLU, :519
    `h + feed_forward(ffn_norm(h))`; the add of `out` onto the stream is the next op's fused residual add).
    Returns (out [1, hidden] fp16, residual_out [1, hidden] fp16 = fp16(x + residual))."""
    ffn = w_down_t.shape[0]
    h = x.reshape(-1).float() + residual.reshape(-1).float()
    res_out = h.half()
    h = res_out.float()          # the residual stream is an fp16 tensor in the model, in either flavour
    n = rmsnorm(h, rms_w, eps, mode)
    gu = w_gate_up.float() @ n
    g, u = _r16(gu[:ffn], mode), _r16(gu[ffn:], mode)
    a = _r16(_r16(torch.nn.functional.silu(g), mode) * u, mode)
    out = w_down_t.float().t() @ a
    return out.half().view(1, -1), res_out.view(1, -1)


# --------------------------------------------------------------------------------------
# the eager fp16 CPU layer, native half tensors -- the timed CPU baseline
# --------------------------------------------------------------------------------------
def eager_fp16_cpu_layer(x, wq, wk, wv, wo, cache_k, cache_v, rms_w, freqs_cis, pos, eps=1e-5):
    """The reference's eager decode step for one layer's attention half, on CPU, in
    native fp16 tensors, written the way chat/llama/model.py writes it
    (:54-79 norm, :376-405 projections/cache update/attention, :134-163 RoPE).
    ``cache_k/v`` [max_seq, Hkv, D] are updated in place at ``pos``.
    Used by bench.py as the CPU baseline (kind="port"): the reference module itself
    cannot be imported without fairscale / fire / a CUDA flashinfer.
    """
    import torch.nn.functional as F
    hq = wq.shape[0] // HEAD_DIM
    hkv = wk.shape[0] // HEAD_DIM
    xf = x.float()
    n = (xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)).type_as(x) * rms_w
    xq = F.linear(n, wq).view(1, 1, hq, HEAD_DIM)
    xk = F.linear(n, wk).view(1, 1, hkv, HEAD_DIM)
    xv = F.linear(n, wv).view(1, 1, hkv, HEAD_DIM)
    xq_ = torch.view_as_complex(xq.float().reshape(1, 1, hq, -1, 2))
    xk_ = torch.view_as_complex(xk.float().reshape(1, 1, hkv, -1, 2))
    fc = freqs_cis[pos:pos + 1].view(1, 1, 1, -1)
    xq = torch.view_as_real(xq_ * fc).flatten(3).type_as(x)
    xk = torch.view_as_real(xk_ * fc).flatten(3).type_as(x)
    cache_k[pos] = xk[0, 0]
    cache_v[pos] = xv[0, 0]
    keys = cache_k[: pos + 1]
    values = cache_v[: pos + 1]
    rep = hq // hkv
    if rep > 1:
        keys = keys[:, :, None, :].expand(-1, hkv, rep, -1).reshape(pos + 1, hq, HEAD_DIM)
        values = values[:, :, None, :].expand(-1, hkv, rep, -1).reshape(pos + 1, hq, HEAD_DIM)
    q = xq.view(hq, 1, HEAD_DIM)
    scores = torch.matmul(q, keys.permute(1, 2, 0)) / math.sqrt(HEAD_DIM)
    scores = F.softmax(scores.float(), dim=-1).type_as(x)
    o = torch.matmul(scores, values.permute(1, 0, 2)).reshape(1, hq * HEAD_DIM)
    return x + F.linear(o, wo)


# --------------------------------------------------------------------------------------
# synthetic inputs, shared by tests / bench / smoke
# --------------------------------------------------------------------------------------
def rope_angles(pos: int, head_dim: int = HEAD_DIM, theta: float = 10000.0) -> torch.Tensor:
    """chat/llama/model.py:82-106 ``precompute_freqs_cis`` at one position -> [D/2] angles."""
    freqs = 1.0 / (theta ** (torch.arange(0, head_dim, 2)[: head_dim // 2].float() / head_dim))
    return (float(pos) * freqs).float()


def make_inputs(shape: LayerShape, kv_len: int, seed: int = 42, w_scale: float = 0.02,
                layout: str = "sglang", theta: float = 10000.0, bs: int = 1):
    """Deterministic synthetic tensors (CPU, fp16) of the named shape.

    layout="chat": W^T-stacked weights (MHA only), cos/sin [1, D] pair-repeated.
    layout="sglang": nn.Linear weights, cos/sin [D/2].
    """
    g = torch.Generator().manual_seed(seed)
    H, D = shape.hidden, shape.head_dim

    def rn(*s, scale=1.0):
        return (torch.randn(*s, generator=g, dtype=torch.float32) * scale).half()

    d = {}
    d["x"] = rn(bs, H)
    d["residual"] = rn(bs, H)
    wq = rn(shape.q_dim, H, scale=w_scale)
    wk = rn(shape.kv_dim, H, scale=w_scale)
    wv = rn(shape.kv_dim, H, scale=w_scale)
    wo = rn(H, shape.q_dim, scale=w_scale)
    d["rms_w"] = (1.0 + 0.1 * torch.randn(H, generator=g)).half()
    d["k_cache"] = rn(kv_len, shape.kv_dim)
    d["v_cache"] = rn(kv_len, shape.kv_dim)
    ang = rope_angles(kv_len, D, theta)
    if layout == "chat":
        assert shape.n_heads == shape.n_kv_heads
        d["weight_qkv"] = torch.cat([wq.t(), wk.t(), wv.t()], 0).contiguous()
        d["weight_o"] = wo.t().contiguous()
        d["cos"] = torch.repeat_interleave(ang.cos(), 2).view(1, D).contiguous()
        d["sin"] = torch.repeat_interleave(ang.sin(), 2).view(1, D).contiguous()
    else:
        d["weight_qkv"] = torch.cat([wq, wk, wv], 0).contiguous()
        d["weight_o"] = wo.contiguous()
        d["cos"] = ang.cos().contiguous()
        d["sin"] = ang.sin().contiguous()
    return d

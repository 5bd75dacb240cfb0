"""Mint golden vectors for the oracle by EXECUTING THE REFERENCE'S OWN PYTHON.

TEST INFRASTRUCTURE ONLY (see oracle/llama_oracle.py header).

Run in the build container, where /root/reference is mounted:

    python -m oracle.gen_golden            # writes tests/golden/*.npz

The GPU box has no /root/reference, so nothing at test / bench time runs this
script; the committed .npz files travel instead.

What is executed, unmodified, from /root/reference:

1. ``tests/test_llama_tilelang.py::reference`` (:18-49) -- the reference's
   pure-torch statement of the sglang signature (fp32 arithmetic).
2. ``chat/llama/model.py``: ``RMSNorm`` (:36-79), ``precompute_freqs_cis``
   (:82-106), ``apply_rotary_emb`` (:134-163), ``Attention.__init__/forward``
   eager branch (:180-270, :376-405) and ``Attention._build_cf_weights``
   (:292-328, the W^T fused-weight layout the 8-arg kernel consumes), with the
   packages that are absent from this image replaced by minimal stand-ins:
     * fairscale Column/RowParallelLinear -> ``nn.Linear`` (world size 1:
       their arithmetic is F.linear);
     * ``flashinfer.single_decode_with_kv_cache`` -> softmax(q K^T / sqrt(d)) V
       in fp32 (its published definition; unpinned by the reference);
     * the native ``clusterfusion`` import -> empty module;
     * ``Tensor.cuda()`` -> identity (no GPU here).

Inputs are NOT stored (a 7B layer's weights are 128 MiB): they are regenerated
from ``oracle.llama_oracle.make_inputs(seed)`` and their SHA-256 is stored next
to the expected outputs, so a drift in the generator is detected, not absorbed.
"""
from __future__ import annotations

import hashlib
import importlib.util
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch
from torch import nn

from oracle.llama_oracle import LayerShape, make_inputs

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"


def sha(t: torch.Tensor) -> str:
    return hashlib.sha256(t.contiguous().view(torch.uint8).numpy().tobytes()).hexdigest()


def inputs_digest(d: dict) -> str:
    h = hashlib.sha256()
    for k in sorted(d):
        h.update(k.encode())
        h.update(sha(d[k]).encode())
    return h.hexdigest()


# --------------------------------------------------------------------------------------
def load_tilelang_reference():
    spec = importlib.util.spec_from_file_location(
        "ref_test_llama_tilelang", REF / "tests" / "test_llama_tilelang.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)          # module level imports only math + torch
    return m.reference


def load_reference_model():
    """Import /root/reference/chat/llama/model.py with absent third parties stubbed."""
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m

    fs = mod("fairscale")
    fs.nn = mod("fairscale.nn")
    fs.nn.model_parallel = mod("fairscale.nn.model_parallel")
    init = mod("fairscale.nn.model_parallel.initialize")
    init.get_model_parallel_world_size = lambda: 1
    fs.nn.model_parallel.initialize = init
    layers = mod("fairscale.nn.model_parallel.layers")

    class ColumnParallelLinear(nn.Linear):
        def __init__(self, i, o, bias=False, gather_output=False, init_method=None):
            super().__init__(i, o, bias=bias)

    class RowParallelLinear(nn.Linear):
        def __init__(self, i, o, bias=False, input_is_parallel=True, init_method=None):
            super().__init__(i, o, bias=bias)

    class ParallelEmbedding(nn.Embedding):
        def __init__(self, n, d, init_method=None):
            super().__init__(n, d)

    layers.ColumnParallelLinear = ColumnParallelLinear
    layers.RowParallelLinear = RowParallelLinear
    layers.ParallelEmbedding = ParallelEmbedding

    cf = mod("clusterfusion")
    cf.llama_decoder_layer = None

    fi = mod("flashinfer")

    def single_decode_with_kv_cache(q, k, v, layout="NHD", pos="NONE", use_tensor_cores=False):
        # q [H, D]; k, v [S, H, D]  (NHD).  Published definition of the op.
        assert layout == "NHD" and pos == "NONE"
        d = q.shape[-1]
        s = torch.einsum("hd,shd->hs", q.float(), k.float()) / (d ** 0.5)
        p = torch.softmax(s, dim=-1)
        return torch.einsum("hs,shd->hd", p, v.float()).to(q.dtype)

    fi.single_decode_with_kv_cache = single_decode_with_kv_cache

    torch.Tensor.cuda = lambda self, *a, **k: self      # no GPU in the build container

    os.environ["USE_CLUSTER_FUSION"] = "true"           # builds rotary buffers + _build_cf_weights
    spec = importlib.util.spec_from_file_location("ref_llama_model", REF / "chat" / "llama" / "model.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


# --------------------------------------------------------------------------------------
def gen_sglang(reference, kv_len: int, seed: int, w_scale: float):
    shape = LayerShape(4096, 32, 32)
    d = make_inputs(shape, kv_len, seed=seed, w_scale=w_scale, layout="sglang")
    eps = 1e-5
    out, res, k, v = reference(d["x"], d["residual"], d["weight_qkv"], d["weight_o"],
                               d["k_cache"], d["v_cache"], d["rms_w"], eps, d["cos"], d["sin"])
    return dict(kind="sglang", kv_len=kv_len, seed=seed, w_scale=w_scale, eps=eps,
                hidden=4096, n_heads=32, n_kv_heads=32, digest=inputs_digest(d),
                out=out.numpy(), residual_out=res.numpy(), k=k.numpy(), v=v.numpy())


def gen_chat(model, shape: LayerShape, kv_len: int, seed: int, w_scale: float, dtype):
    """Reference eager Attention on CPU; returns outputs + the reference-built fused weights' hash."""
    d = make_inputs(shape, kv_len, seed=seed, w_scale=w_scale, layout="sglang")
    eps = 1e-6
    args = model.ModelArgs(dim=shape.hidden, n_heads=shape.n_heads,
                           n_kv_heads=shape.n_kv_heads, max_batch_size=1,
                           max_seq_len=max(kv_len + 1, 2))
    torch.set_default_dtype(dtype)
    try:
        attn = model.Attention(args)
        norm = model.RMSNorm(shape.hidden, eps=eps)
    finally:
        torch.set_default_dtype(torch.float32)
    wq, wk, wv = d["weight_qkv"].split([shape.q_dim, shape.kv_dim, shape.kv_dim], 0)
    with torch.no_grad():
        attn.wq.weight.copy_(wq)
        attn.wk.weight.copy_(wk)
        attn.wv.weight.copy_(wv)
        attn.wo.weight.copy_(d["weight_o"])
        norm.weight.copy_(d["rms_w"])
        attn.cache_k = attn.cache_k.to(dtype)
        attn.cache_v = attn.cache_v.to(dtype)
        attn.cache_k[0, :kv_len] = d["k_cache"].view(kv_len, shape.n_kv_heads, shape.head_dim)
        attn.cache_v[0, :kv_len] = d["v_cache"].view(kv_len, shape.n_kv_heads, shape.head_dim)
        # the fused weights exactly as the reference lays them out for the 8-arg kernel
        fused = {}
        if shape.n_heads == shape.n_kv_heads:
            attn._build_cf_weights()
            fused = dict(weight_qkv_sha=sha(attn.weight_qkv.half()), weight_o_sha=sha(attn.weight_o.half()),
                         cos=attn.rotary_cos[kv_len:kv_len + 1].float().numpy(),
                         sin=attn.rotary_sin[kv_len:kv_len + 1].float().numpy())
        freqs_cis = model.precompute_freqs_cis(shape.head_dim, args.max_seq_len * 2)[kv_len:kv_len + 1]
        x = d["x"].to(dtype).view(1, 1, shape.hidden)
        attn.use_cluster_fusion = False                 # force the eager branch (:376-405)
        o = attn.forward(norm(x), kv_len, freqs_cis, None, norm.weight)
        k_new = attn.cache_k[0, kv_len].clone()
        v_new = attn.cache_v[0, kv_len].clone()
    return dict(kind="chat", dtype=str(dtype).split(".")[-1], kv_len=kv_len, seed=seed, w_scale=w_scale,
                eps=eps, hidden=shape.hidden, n_heads=shape.n_heads, n_kv_heads=shape.n_kv_heads,
                digest=inputs_digest(d),
                out=o.reshape(1, -1).float().numpy().astype(np.float32),
                k=k_new.float().numpy().astype(np.float32)[None],
                v=v_new.float().numpy().astype(np.float32)[None], **fused)


def gen_ffn(model, seed: int, dtype):
    """Reference FeedForward (chat/llama/model.py:407-448) + RMSNorm (:36-79) on CPU, Llama-2-7B dims (ffn 11008)."""
    hidden = 4096
    g = torch.Generator().manual_seed(seed)
    rn = lambda *sh, sc=1.0: (torch.randn(*sh, generator=g, dtype=torch.float32) * sc).half()
    torch.set_default_dtype(dtype)
    try:
        ff = model.FeedForward(dim=hidden, hidden_dim=4 * hidden, multiple_of=256, ffn_dim_multiplier=None)
        norm = model.RMSNorm(hidden, eps=1e-5)
    finally:
        torch.set_default_dtype(torch.float32)
    ffn = ff.w1.weight.shape[0]
    x, residual = rn(1, hidden), rn(1, hidden)
    w1, w3, w2 = rn(ffn, hidden, sc=0.02), rn(ffn, hidden, sc=0.02), rn(hidden, ffn, sc=0.02)
    rms = (1.0 + 0.1 * torch.randn(hidden, generator=g)).half()
    with torch.no_grad():
        ff.w1.weight.copy_(w1); ff.w3.weight.copy_(w3); ff.w2.weight.copy_(w2); norm.weight.copy_(rms)
        h = (x.float() + residual.float()).half().to(dtype)            # the stream is an fp16 tensor in the model
        out = ff(norm(h.view(1, 1, hidden)))
    d = dict(x=x, residual=residual, w1=w1, w3=w3, w2=w2, rms=rms)
    return dict(kind="ffn", dtype=str(dtype).split(".")[-1], seed=seed, hidden=hidden, ffn=ffn, eps=1e-5,
                digest=inputs_digest(d), out=out.reshape(1, -1).float().numpy().astype(np.float32))


def ffn_inputs(seed: int, hidden: int = 4096, ffn: int = 11008):
    """Regenerate the tensors gen_ffn drew (same generator order)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *sh, sc=1.0: (torch.randn(*sh, generator=g, dtype=torch.float32) * sc).half()
    x, residual = rn(1, hidden), rn(1, hidden)
    w1, w3, w2 = rn(ffn, hidden, sc=0.02), rn(ffn, hidden, sc=0.02), rn(hidden, ffn, sc=0.02)
    rms = (1.0 + 0.1 * torch.randn(hidden, generator=g)).half()
    return dict(x=x, residual=residual, w1=w1, w3=w3, w2=w2, rms=rms)


def main():
    assert REF.exists(), "run in the build container (needs /root/reference)"
    OUT.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    reference = load_tilelang_reference()
    model = load_reference_model()

    cases = []
    for kv_len in (0, 1, 37, 256, 1024):
        cases.append((f"sglang_kv{kv_len}_w0.02", lambda kv=kv_len: gen_sglang(reference, kv, 42, 0.02)))
    # the reference's own test distribution: weights * 0.1 (tests/test_llama_tilelang.py:81-82)
    cases.append(("sglang_kv37_w0.1", lambda: gen_sglang(reference, 37, 7, 0.1)))
    s7 = LayerShape(4096, 32, 32)
    for kv_len in (0, 1, 37, 256, 1024):
        cases.append((f"chat_fp32_kv{kv_len}", lambda kv=kv_len: gen_chat(model, s7, kv, 42, 0.02, torch.float32)))
    for kv_len in (37, 1024):
        cases.append((f"chat_fp16_kv{kv_len}", lambda kv=kv_len: gen_chat(model, s7, kv, 42, 0.02, torch.float16)))
    # GQA through the reference's eager Attention (repeat_kv, model.py:166-175)
    cases.append(("chat_fp32_gqa8_kv100", lambda: gen_chat(model, LayerShape(4096, 32, 8), 100, 11, 0.02, torch.float32)))
    cases.append(("chat_fp32_70b_kv64", lambda: gen_chat(model, LayerShape(8192, 64, 8), 64, 13, 0.02, torch.float32)))

    cases.append(("ffn_fp32_seed21", lambda: gen_ffn(model, 21, torch.float32)))
    cases.append(("ffn_fp16_seed21", lambda: gen_ffn(model, 21, torch.float16)))
    only = os.environ.get("GOLDEN_ONLY")
    if only:
        cases = [c for c in cases if only in c[0]]

    for name, fn in cases:
        r = fn()
        np.savez_compressed(OUT / f"{name}.npz", **r)
        print(f"{name}: out|max|={np.abs(r['out']).max():.4f} digest={r['digest'][:12]}")


if __name__ == "__main__":
    main()

"""Pin the CPU oracle against fixtures produced by executing the reference's own Python
(oracle/gen_golden.py; reference tests/test_llama_tilelang.py:18-49 and
chat/llama/model.py eager Attention).  CPU only."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import llama_oracle as O
from oracle.gen_golden import inputs_digest, sha

from conftest import GOLDEN

SGLANG = sorted(p.name for p in GOLDEN.glob("sglang_*.npz"))
CHAT = sorted(p.name for p in GOLDEN.glob("chat_*.npz"))


def _load(name):
    z = np.load(GOLDEN / name, allow_pickle=False)
    return {k: z[k] for k in z.files}


def _maxdiff(a, b):
    return float((torch.as_tensor(np.asarray(a, dtype=np.float32)) - torch.as_tensor(np.asarray(b, dtype=np.float32))).abs().max())


def test_fixtures_present():
    assert len(SGLANG) >= 6 and len(CHAT) >= 9


@pytest.mark.parametrize("name", SGLANG)
def test_sglang_oracle_matches_reference_python(name):
    g = _load(name)
    shape = O.LayerShape(int(g["hidden"]), int(g["n_heads"]), int(g["n_kv_heads"]))
    d = O.make_inputs(shape, int(g["kv_len"]), seed=int(g["seed"]), w_scale=float(g["w_scale"]), layout="sglang")
    assert inputs_digest(d) == str(g["digest"]), "synthetic-input generator drifted from the fixture"
    out, res, k, v = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"],
                                    d["v_cache"], d["rms_w"], float(g["eps"]), d["cos"], d["sin"], mode="fp32")
    # identical algorithm in fp32 on the same inputs: differences are summation order only.
    scale = max(1.0, float(np.abs(g["out"].astype(np.float32)).max()))
    assert _maxdiff(out, g["out"]) <= 1e-3 * scale       # <= 1 fp16 ulp of the largest element
    assert torch.equal(res, torch.from_numpy(g["residual_out"]))
    assert _maxdiff(k, g["k"]) <= 4e-3 and _maxdiff(v, g["v"]) <= 4e-3
    # and the eager-rounding flavour stays inside the north-star tolerance of the fp32 flavour
    out_e, _, k_e, v_e = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"],
                                        d["v_cache"], d["rms_w"], float(g["eps"]), d["cos"], d["sin"], mode="eager")
    if float(g["w_scale"]) <= 0.02:
        # with almost no keys to average over (kv < 37) the fp16 rounding of v_new shows 1:1 in the output
        tol = 1e-3 if int(g["kv_len"]) >= 37 else 2.5e-3
        assert torch.allclose(out_e.float(), torch.from_numpy(g["out"]).float(), rtol=tol, atol=tol)


@pytest.mark.parametrize("name", CHAT)
def test_chat_oracle_matches_reference_eager_attention(name):
    g = _load(name)
    shape = O.LayerShape(int(g["hidden"]), int(g["n_heads"]), int(g["n_kv_heads"]))
    kv = int(g["kv_len"])
    d = O.make_inputs(shape, kv, seed=int(g["seed"]), w_scale=float(g["w_scale"]), layout="sglang")
    assert inputs_digest(d) == str(g["digest"])
    wq, wk, wv = d["weight_qkv"].split([shape.q_dim, shape.kv_dim, shape.kv_dim], 0)
    ang = O.rope_angles(kv, shape.head_dim)
    cos = torch.repeat_interleave(ang.cos(), 2).view(1, -1)
    sin = torch.repeat_interleave(ang.sin(), 2).view(1, -1)
    mode = "eager" if str(g["dtype"]) == "float16" else "fp32"
    if shape.n_heads == shape.n_kv_heads:
        # the 8-arg layout: must be byte-identical to what the reference's _build_cf_weights made
        wqkv_T = torch.cat([wq.t(), wk.t(), wv.t()], 0).contiguous()
        wo_T = d["weight_o"].t().contiguous()
        assert sha(wqkv_T) == str(g["weight_qkv_sha"])
        assert sha(wo_T) == str(g["weight_o_sha"])
        assert _maxdiff(cos, g["cos"]) < 1e-6 and _maxdiff(sin, g["sin"]) < 1e-6
        out, k, v = O.chat_layer(d["x"], wqkv_T, wo_T, d["k_cache"], d["v_cache"], d["rms_w"], cos, sin,
                                 n_heads=shape.n_heads, eps=float(g["eps"]), mode=mode)
        out, k, v = out.float(), k.float(), v.float()
    else:
        o32, k, v = O._core(d["x"].reshape(-1).float(), wq, wk, wv, d["weight_o"], d["k_cache"], d["v_cache"],
                            d["rms_w"], float(g["eps"]), cos, sin, "gptj", shape, mode)
        out = o32.view(1, -1)
        k, v = k[None], v[None]
    ref_out = torch.from_numpy(g["out"].astype(np.float32))
    if mode == "fp32":
        # the fixture is unrounded fp32; the oracle's public output is fp16-rounded
        assert torch.allclose(out, ref_out, rtol=1e-3, atol=1e-3)
        assert _maxdiff(out, ref_out) <= 1e-3 * max(1.0, float(ref_out.abs().max()))
        assert _maxdiff(k.reshape(-1), g["k"].reshape(-1)) <= 4e-3
        assert _maxdiff(v.reshape(-1), g["v"].reshape(-1)) <= 2e-3
    else:
        # reference ran natively in fp16 on CPU: same rounding points, <= ~2 ulp apart
        assert torch.allclose(out, ref_out, rtol=2e-3, atol=2e-3)
        assert _maxdiff(k.reshape(-1), g["k"].reshape(-1)) <= 8e-3
        assert _maxdiff(v.reshape(-1), g["v"].reshape(-1)) <= 4e-3


def test_chat_and_sglang_agree_under_relayout():
    """Same maths, two layouts: W^T + GPT-J pairs  ==  [out,in] + NeoX after permuting head dims."""
    shape = O.LayerShape(1024, 8, 8)
    d = O.make_inputs(shape, 19, seed=3, layout="sglang")
    wq, wk, wv = d["weight_qkv"].split([1024, 1024, 1024], 0)
    D = 128
    # NeoX index j<64 pairs (j, j+64); GPT-J pairs (2j, 2j+1): permute rows of Wq/Wk within each head
    perm = torch.empty(D, dtype=torch.long)
    perm[0::2] = torch.arange(0, D // 2)
    perm[1::2] = torch.arange(D // 2, D)
    def to_gptj(w):
        return w.view(8, D, -1)[:, perm, :].reshape(1024, -1)
    wqkv_T = torch.cat([to_gptj(wq).t(), to_gptj(wk).t(), wv.t()], 0).contiguous()
    ang = O.rope_angles(19)
    cos = torch.repeat_interleave(ang.cos(), 2).view(1, -1)
    sin = torch.repeat_interleave(ang.sin(), 2).view(1, -1)
    o_c, k_c, v_c = O.chat_layer(d["x"], wqkv_T, d["weight_o"].t().contiguous(), d["k_cache"].view(19, 8, D)[:, :, perm].reshape(19, -1),
                                 d["v_cache"], d["rms_w"], cos, sin, n_heads=8, eps=1e-5)
    zero = torch.zeros_like(d["x"])
    o_s, _, k_s, v_s = O.sglang_layer(d["x"], zero, d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                                      d["rms_w"], 1e-5, ang.cos(), ang.sin(), n_heads=8)
    assert torch.allclose(o_c.float(), o_s.float(), rtol=1e-3, atol=1e-3)
    assert torch.allclose(k_c.float()[0], k_s.float()[0][:, perm], rtol=1e-3, atol=2e-3)
    assert torch.allclose(v_c.float(), v_s.float(), rtol=1e-3, atol=1e-3)   # summation order differs


def test_paged_equals_contiguous():
    shape = O.LayerShape(1024, 8, 8)
    bs, lens = 3, [5, 0, 33]
    d = O.make_inputs(shape, 64, seed=5, layout="sglang", bs=bs)
    g = torch.Generator().manual_seed(0)
    slots = torch.randperm(64, generator=g)
    indptr, indices, off = [0], [], 0
    for L in lens:
        indices += slots[off:off + L + 1].tolist()
        off += L + 1
        indptr.append(len(indices))
    indptr = torch.tensor(indptr, dtype=torch.int32)
    indices = torch.tensor(indices, dtype=torch.int32)
    positions = torch.tensor(lens, dtype=torch.int64)
    cos_sin = torch.stack([torch.cat([O.rope_angles(p).cos(), O.rope_angles(p).sin()]) for p in range(40)])
    kp, vp = d["k_cache"].clone(), d["v_cache"].clone()
    out, res = O.paged_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], indptr, indices, kp, vp,
                             d["rms_w"], 1e-5, positions, cos_sin, n_heads=8)
    for b, L in enumerate(lens):
        rows = indices[indptr[b]:indptr[b + 1] - 1].long()
        a = O.rope_angles(L)
        o, r, k, v = O.sglang_layer(d["x"][b:b + 1], d["residual"][b:b + 1], d["weight_qkv"], d["weight_o"],
                                    d["k_cache"][rows], d["v_cache"][rows], d["rms_w"], 1e-5, a.cos(), a.sin(), n_heads=8)
        assert torch.equal(out[b:b + 1], o) and torch.equal(res[b:b + 1], r)
        slot = int(indices[indptr[b + 1] - 1])
        assert torch.equal(kp[slot], k.reshape(-1)) and torch.equal(vp[slot], v.reshape(-1))
    untouched = [i for i in range(64) if i not in {int(indices[indptr[b + 1] - 1]) for b in range(bs)}]
    assert torch.equal(kp[untouched], d["k_cache"][untouched])


def test_eager_fp16_cpu_layer_matches_oracle():
    """The timed CPU baseline (native fp16 eager ops) computes the same thing as the oracle."""
    shape = O.LayerShape(1024, 8, 8)
    kv = 50
    d = O.make_inputs(shape, kv, seed=9, layout="sglang")
    wq, wk, wv = d["weight_qkv"].split([1024, 1024, 1024], 0)
    ck = torch.zeros(kv + 1, 8, 128, dtype=torch.float16)
    cv = torch.zeros(kv + 1, 8, 128, dtype=torch.float16)
    ck[:kv] = d["k_cache"].view(kv, 8, 128)
    cv[:kv] = d["v_cache"].view(kv, 8, 128)
    ang = torch.stack([O.rope_angles(p) for p in range(kv + 1)])
    fc = torch.polar(torch.ones_like(ang), ang)
    y = O.eager_fp16_cpu_layer(d["x"], wq, wk, wv, d["weight_o"], ck, cv, d["rms_w"], fc, kv, eps=1e-6)
    a = O.rope_angles(kv)
    cos = torch.repeat_interleave(a.cos(), 2).view(1, -1)
    sin = torch.repeat_interleave(a.sin(), 2).view(1, -1)
    wqkv_T = torch.cat([wq.t(), wk.t(), wv.t()], 0).contiguous()
    o, k, v = O.chat_layer(d["x"], wqkv_T, d["weight_o"].t().contiguous(), d["k_cache"], d["v_cache"], d["rms_w"],
                           cos, sin, n_heads=8, eps=1e-6, mode="eager")
    assert torch.allclose((y - d["x"]).float(), o.float(), rtol=5e-3, atol=5e-3)
    assert torch.allclose(ck[kv].float(), k[0].float(), rtol=2e-3, atol=4e-3)


@pytest.mark.parametrize("name", sorted(p.name for p in GOLDEN.glob("ffn_*.npz")))
def test_ffn_oracle_matches_reference_feedforward(name):
    """oracle.ffn_layer against the reference's FeedForward + RMSNorm modules run on CPU (chat/llama/model.py:407-448)."""
    from oracle.gen_golden import ffn_inputs
    g = _load(name)
    d = ffn_inputs(int(g["seed"]))
    assert inputs_digest(d) == str(g["digest"])
    mode = "eager" if str(g["dtype"]) == "float16" else "fp32"
    w13 = torch.cat([d["w1"], d["w3"]], 0).contiguous()
    w2t = d["w2"].t().contiguous()
    out, res = O.ffn_layer(d["x"], d["residual"], w13, w2t, d["rms"], float(g["eps"]), mode=mode)
    ref = torch.from_numpy(g["out"])
    assert torch.allclose(out.float(), ref, rtol=1e-3, atol=1e-3)      # same flavour as the reference run
    assert torch.equal(res, (d["x"].float() + d["residual"].float()).half())
    # the two flavours agree to the north-star tolerance
    out2, _ = O.ffn_layer(d["x"], d["residual"], w13, w2t, d["rms"], float(g["eps"]), mode="eager" if mode == "fp32" else "fp32")
    # across flavours the fp16 rounding of the 11008 activations shows: |out| ~ 2, a few 1e-3 apart
    assert torch.allclose(out.float(), out2.float(), rtol=4e-3, atol=4e-3)


def test_rmsnorm_op_oracle_definition():
    """Standalone op (row f4): fp32 math, one rounding; rows are independent; weight == 1 and unit-RMS rows are fixed points."""
    import torch
    from oracle import llama_oracle as O
    g = torch.Generator().manual_seed(1)
    x = torch.randn(5, 512, generator=g).half()
    w = torch.randn(512, generator=g).half()
    y = O.rmsnorm_op(x, w)
    assert y.dtype == torch.float16 and y.shape == x.shape
    for b in range(5):
        assert torch.equal(y[b:b + 1], O.rmsnorm_op(x[b:b + 1], w))
    xf = x.double()
    want = (xf / torch.sqrt(xf.pow(2).mean(-1, keepdim=True) + 1e-6) * w.double()).half()
    assert float((y.float() - want.float()).abs().max()) <= float(want.float().abs().max()) * 2 ** -10
    ones = torch.ones(512).half()
    unit = torch.ones(1, 512).half()
    assert torch.equal(O.rmsnorm_op(unit, ones), unit)

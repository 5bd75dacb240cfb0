"""CPU checks of the DeepSeek-MLA oracle (oracle/deepseek_oracle.py).  The reference holds no test or fixture for this op; the
pin is the output of the reference's OWN kernel (oracle/_ref) at the one shape its binary supports, captured on B200 under
compute-sanitizer memcheck, where its shared-memory races do not fire (tests/golden/deepseek_ref_kernel_seq4096.npz, minted by
oracle/gen_golden_deepseek_ref.py).  The rest is internal: closed forms at seq_len 1, the rounding noise of
the fp16 flavour against float64, and the algebraic properties the GPU tests rely on."""
import math

import torch

from oracle import deepseek_oracle as D

ORACLE_DIGEST = "4c25e94b816e3f847d4f8b2e02c23396b014706553f3031930a4d47aacf8e822"


def test_oracle_matches_the_reference_kernels_race_free_output():
    """The fixture is what /root/reference/include/H100/deepseek/kernel.cuh computed on B200 (recompiled unmodified for sm_100a) for the
    seeded inputs below.  The reference sums in fp16 (partials rounded before its cluster reduction, fp16 atomics for the
    O projection, two launches agree to 2e-3), so the comparison is at 1e-2 on outputs of magnitude 1.8 -- five times tighter than
    the only tolerance the reference itself asserts anywhere (5e-2, tests/test_llama_tilelang.py:100)."""
    import numpy as np
    from conftest import GOLDEN
    from oracle.gen_golden_deepseek_ref import KEYS, inputs_digest
    z = np.load(GOLDEN / "deepseek_ref_kernel_seq4096.npz")
    d = D.make_inputs(int(z["seq_len"]), seed=int(z["seed"]), out_gain=float(z["out_gain"]))
    assert inputs_digest(d) == str(z["inputs_sha256"]), "the seeded inputs no longer reproduce the fixture's inputs"
    want = torch.from_numpy(z["out"]).float()
    got, _, _ = D.deepseek_layer(**d)
    err = float((got.float().reshape(-1) - want).abs().max())
    print(f"oracle vs reference kernel: max |diff| {err:.4f} on |out| max {float(want.abs().max()):.3f}")
    assert float(want.abs().max()) > 1.0 and err <= 1e-2
    assert float((torch.from_numpy(z["out_second_launch"]).float() - want).abs().max()) <= 5e-3     # the witness is stable


def test_seq_len_1_closed_form():
    """With no cache rows the softmax has one term: out = W_o (W_uv ckv_n), independent of q, RoPE and the flag."""
    d = D.make_inputs(1, seed=5, out_gain=0.75)
    out, ckv_n, k_pe = D.deepseek_layer(**d, mode="exact")
    x = d["x"].double().view(-1)
    xn = x * torch.rsqrt((x * x).mean() + D.EPS) * d["rms_in_w"].double()
    ckv = xn @ d["w_kv"].double()
    want_ckv = ckv * torch.rsqrt((ckv * ckv).mean() + D.EPS) * d["rms_ckv_w"].double()
    assert torch.allclose(ckv_n.double(), want_ckv, atol=2e-3)
    attn = torch.einsum("k,khn->hn", want_ckv, d["w_uv"].double().view(D.LORA, D.N_HEADS, D.NOPE))
    want = attn.reshape(1, -1) @ d["w_o"].double()
    assert torch.allclose(out.double(), want, rtol=1e-3, atol=1e-3)
    out_r, _, _ = D.deepseek_layer(**d, mode="exact", rope_scores=True)
    assert torch.equal(out, out_r)


def test_eager_flavour_is_rounding_close_to_float64():
    for S, gain in ((2, 0.75), (300, 1.5), (4096, 2.4)):
        d = D.make_inputs(S, seed=S, out_gain=gain)
        for rope in (False, True):
            o, c, k = D.deepseek_layer(**d, rope_scores=rope)
            oe, ce, ke = D.deepseek_layer(**d, rope_scores=rope, mode="exact")
            assert float(o.float().abs().max()) > 0.5                       # the comparison is not vacuous
            assert torch.allclose(o.float(), oe.float(), rtol=2e-3, atol=2e-3), (S, rope)
            assert torch.allclose(c.float(), ce.float(), rtol=2e-3, atol=2e-3)
            assert torch.allclose(k.float(), ke.float(), rtol=2e-3, atol=2e-3)


def test_rope_term_and_row_permutation():
    d = D.make_inputs(200, seed=9, out_gain=1.5)
    o, _, _ = D.deepseek_layer(**d)
    o_r, _, _ = D.deepseek_layer(**d, rope_scores=True)
    assert float((o.float() - o_r.float()).abs().max()) > 0.05              # the flag changes the scores
    # the reference kernel's scores never see cache columns 512..575 (kernel.cuh:400-470 loads latent columns only)
    d2 = dict(d, ckv_cache=d["ckv_cache"].clone())
    d2["ckv_cache"][:, D.LORA:] = 7.0
    assert torch.equal(D.deepseek_layer(**d2)[0], o)
    # the cache's last row is replaced by the current token (kernel.cuh:469-470)
    d2["ckv_cache"][-1] = -3.0
    assert torch.equal(D.deepseek_layer(**d2)[0], o)
    # attention is a set function of the cache rows
    perm = torch.randperm(199, generator=torch.Generator().manual_seed(1))
    d3 = dict(d, ckv_cache=torch.cat([d["ckv_cache"][:199][perm], d["ckv_cache"][199:]]))
    assert torch.allclose(D.deepseek_layer(**d3, rope_scores=True)[0].float(), o_r.float(), rtol=1e-3, atol=1e-3)


def test_rope_matches_complex_rotation():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(3, D.ROPE, generator=g)
    ang = torch.rand(D.ROPE // 2, generator=g) * 6
    cos, sin = torch.cat([ang.cos(), ang.cos()]), torch.cat([ang.sin(), ang.sin()])
    got = D._rope(v, cos, sin)
    z = torch.complex(v[:, : D.ROPE // 2], v[:, D.ROPE // 2:]) * torch.polar(torch.ones_like(ang), ang)
    assert torch.allclose(got, torch.cat([z.real, z.imag], dim=-1), atol=1e-5)
    assert math.isclose(float(got.norm()), float(v.norm()), rel_tol=1e-5)


def test_absorbed_form_equals_explicit_per_head_attention():
    """The kernel (and the oracle) work in the latent space: q_lat = q_nope W_uk scores the 512-wide cache rows directly and
    W_uv is applied after the softmax.  The textbook form materialises per-head keys and values,
    k[h, t] = W_uk[h] row_t (128 dims), v[h, t] = row_t W_uv[h], and attends over those.  Both must agree (float64)."""
    S = 50
    d = D.make_inputs(S, seed=13, out_gain=1.0)
    out, ckv_n, k_pe = D.deepseek_layer(**d, rope_scores=True, mode="exact")
    f = lambda t: t.double()
    x = f(d["x"]).view(-1)
    xn = x * torch.rsqrt((x * x).mean() + D.EPS) * f(d["rms_in_w"])
    q_nope = (xn @ f(d["w_q_nope"])).view(D.N_HEADS, D.NOPE)
    q_pe = D._rope((xn @ f(d["w_q_pe"])).view(D.N_HEADS, D.ROPE), f(d["cos"]), f(d["sin"]))
    rows = torch.cat([f(d["ckv_cache"][: S - 1, : D.LORA]), f(ckv_n)[None]], 0)                    # [S, 512]
    pes = torch.cat([f(d["ckv_cache"][: S - 1, D.LORA:]), f(k_pe)[None]], 0)                       # [S, 64]
    w_uk = f(d["w_uk"]).view(D.NOPE, D.N_HEADS, D.LORA)
    w_uv = f(d["w_uv"]).view(D.LORA, D.N_HEADS, D.NOPE)
    k = torch.einsum("tc,dhc->htd", rows, w_uk)                                                    # [heads, S, 128]
    v = torch.einsum("tc,chn->htn", rows, w_uv)                                                    # [heads, S, 128]
    scores = (torch.einsum("hd,htd->ht", q_nope, k) + q_pe @ pes.T) / math.sqrt(D.NOPE + D.ROPE)
    attn = torch.einsum("ht,htn->hn", torch.softmax(scores, -1), v)
    want = attn.reshape(1, -1) @ f(d["w_o"])
    # the oracle's outputs (out, ckv_n, k_pe) are fp16 tensors: one rounding of the inputs to this check and one of the result
    assert torch.allclose(out.double(), want, rtol=2e-3, atol=2e-3)


def test_oracle_digest_is_stable():
    """Drift guard for the oracle itself (NOT a pin to the reference, which has nothing to pin to): SHA-256 of the float64
    flavour's fp16 outputs on seeded inputs, as committed when the CUDA path was validated against it on B200."""
    import hashlib
    h = hashlib.sha256()
    for S, rope in ((1, False), (37, False), (37, True), (300, True)):
        d = D.make_inputs(S, seed=1000 + S, out_gain=1.0)
        for t in D.deepseek_layer(**d, rope_scores=rope, mode="exact"):
            h.update(t.contiguous().view(torch.uint8).numpy().tobytes())
    assert h.hexdigest() == ORACLE_DIGEST, h.hexdigest()

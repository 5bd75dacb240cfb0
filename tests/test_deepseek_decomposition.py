"""CPU model of HOW the DeepSeek-MLA kernels split the work (clusterfusion_b200/csrc/deepseek_mla_kernel.cuh), checked against the
oracle: hidden-row slices summed in fp32 for the projections, 128 cache slices each keeping an un-normalised flash-decode state
built tile by tile (32 rows) with fp16-rounded probabilities, the current token as state number 128, a max-shifted merge, and
latent-row / head slices summed in fp32 for W_uv / W_o.  It pins the arithmetic of the decomposition -- rounding points, the
log2-domain softmax, the merge formula, empty slices -- on the CPU, where the CUDA code itself cannot run."""
import math

import pytest
import torch

from oracle import deepseek_oracle as D

SPLITS, TILE, CLUSTER = 128, 32, 8
LOG2E = 1.4426950408889634


def rh(t):
    return t.half().float()


def model(d, rope_scores):
    f = lambda t: t.float()
    x = f(d["x"]).view(-1)
    xn = rh(x * torch.rsqrt((x * x).mean() + D.EPS) * f(d["rms_in_w"]))
    S = d["ckv_cache"].shape[0]
    # ---- kernel 1: per head, 8 hidden slices of 256 rows summed in fp32 (cluster_reduce), then W_uk by column slices
    def ksplit(w, parts):
        rows = w.shape[0] // parts
        return sum(xn[i * rows:(i + 1) * rows] @ f(w)[i * rows:(i + 1) * rows] for i in range(parts))
    q_nope = rh(ksplit(d["w_q_nope"], CLUSTER)).view(D.N_HEADS, D.NOPE)
    q_pe = rh(ksplit(d["w_q_pe"], CLUSTER)).view(D.N_HEADS, D.ROPE)
    q_lat = torch.einsum("hk,khn->hn", q_nope, f(d["w_uk"]).view(D.NOPE, D.N_HEADS, D.LORA))
    q_pe_rot = D._rope(q_pe, d["cos"], d["sin"]) if rope_scores else torch.zeros_like(q_pe)
    q = torch.cat([q_lat, q_pe_rot], 1).half().float()                                # [16, 576] as stored in the workspace
    # shared projection: 128 slices of 16 hidden rows, fp32 partials added in the workspace
    acc = ksplit(torch.cat([d["w_kv"], d["w_k_pe"]], 1), SPLITS)
    ckv = rh(acc[: D.LORA])
    ckv_n = rh(ckv * torch.rsqrt((ckv * ckv).mean() + D.EPS) * f(d["rms_ckv_w"]))
    k_pe = rh(D._rope(rh(acc[D.LORA:]), d["cos"], d["sin"]))
    tok = torch.cat([ckv_n, k_pe])
    # ---- kernel 2: one flash-decode state per cache slice, tiles of 32 rows, two 16-row groups merged at the end
    n_rows = S - 1
    per = -(-n_rows // SPLITS) if n_rows else 0
    rps = TILE if per <= TILE else -(-per // TILE) * TILE
    scale_log2 = LOG2E / math.sqrt(D.NOPE + D.ROPE)
    cache = f(d["ckv_cache"])
    ms, ls, os_ = [], [], []
    for s in range(SPLITS):
        r0, r1 = s * rps, min(s * rps + rps, n_rows)
        grp = []
        for rg in range(2):                                                           # warp row groups keep separate states
            m = torch.full((D.N_HEADS,), -math.inf)
            l = torch.zeros(D.N_HEADS)
            o = torch.zeros(D.N_HEADS, D.LORA)
            for t0 in range(r0, max(r1, r0), TILE):
                a, b = t0 + rg * 16, min(t0 + rg * 16 + 16, r1)
                if b <= a:
                    continue
                rows = cache[a:b]
                sc = (q @ rows.T) * scale_log2
                m_new = torch.maximum(m, sc.max(1).values)
                corr = torch.where(torch.isinf(m), torch.zeros_like(m), torch.exp2(m - m_new))
                p = rh(torch.exp2(sc - m_new[:, None]))                               # fp16 probabilities feed the PV product AND the row sum
                l = l * corr + p.sum(1)
                o = o * corr[:, None] + p @ rows[:, : D.LORA]
                m = m_new
            grp.append((m, l, o))
        (m1, l1, o1), (m2, l2, o2) = grp
        M = torch.maximum(m1, m2)
        Mu = torch.where(torch.isinf(M), torch.zeros_like(M), M)
        w1 = torch.where(torch.isinf(m1), torch.zeros_like(m1), torch.exp2(m1 - Mu))
        w2 = torch.where(torch.isinf(m2), torch.zeros_like(m2), torch.exp2(m2 - Mu))
        ms.append(M); ls.append(l1 * w1 + l2 * w2); os_.append(o1 * w1[:, None] + o2 * w2[:, None])
    # the current token: state number 128
    ms.append((q @ tok) * scale_log2); ls.append(torch.ones(D.N_HEADS)); os_.append(ckv_n[None].expand(D.N_HEADS, -1))
    ms, ls, os_ = torch.stack(ms), torch.stack(ls), torch.stack(os_)                  # [129, 16], [129, 16], [129, 16, 512]
    # ---- kernel 3: merge, W_uv by 8 latent slices of 64 rows, W_o by heads, all sums fp32
    M = ms.max(0).values
    w = torch.where(torch.isinf(ms), torch.zeros_like(ms), torch.exp2(ms - M))
    o_lat = rh((w[:, :, None] * os_).sum(0) / (w * ls).sum(0)[:, None])
    w_uv = f(d["w_uv"]).view(D.LORA, D.N_HEADS, D.NOPE)
    attn = rh(sum(torch.einsum("hk,khn->hn", o_lat[:, i * 64:(i + 1) * 64], w_uv[i * 64:(i + 1) * 64]) for i in range(CLUSTER)))
    w_o = f(d["w_o"]).view(D.N_HEADS, D.NOPE, D.HIDDEN)
    out = sum(attn[h] @ w_o[h] for h in range(D.N_HEADS))
    return out.half().view(1, -1), ckv_n.half(), k_pe.half()


@pytest.mark.parametrize("rope", [False, True])
@pytest.mark.parametrize("seq_len,gain", [(1, 0.75), (2, 0.75), (34, 1.0), (300, 1.5), (4096, 2.4), (5000, 2.4)])
def test_kernel_decomposition_matches_oracle(seq_len, gain, rope):
    d = D.make_inputs(seq_len, seed=seq_len, out_gain=gain)
    want, ckv, kpe = D.deepseek_layer(**d, rope_scores=rope)
    got, gckv, gkpe = model(d, rope)
    assert torch.allclose(got.float(), want.float(), rtol=1e-3, atol=1e-3), float((got.float() - want.float()).abs().max())
    assert torch.allclose(gckv.float(), ckv.float(), rtol=2e-3, atol=2e-3)
    assert torch.allclose(gkpe.float(), kpe.float(), rtol=2e-3, atol=2e-3)

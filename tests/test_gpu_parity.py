"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the pybind
operators and through the raw C ABI, against the CPU oracle on the same seeded inputs and against
the committed golden fixtures.  Tolerance: rtol = atol = 1e-3 in fp16 (BASELINE.json north_star)."""
import math

import numpy as np
import pytest
import torch

from oracle import llama_oracle as O
from oracle.gen_golden import inputs_digest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

from parity_helpers import ATOL, RTOL, close, close_k  # noqa: E402,F401

S7 = O.LayerShape(4096, 32, 32)


def cuda(d):
    return {k: v.cuda() for k, v in d.items()}


# ---------------------------------------------------------------------------------------------------
# device primitive (include/dsm.cuh) in isolation
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cluster", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("repeats", [1, 5])
def test_cluster_reduce_sum(cluster, repeats):
    from clusterfusion_b200 import cabi
    n, ncl = 384, 3
    g = torch.Generator().manual_seed(cluster * 10 + repeats)
    x = torch.randn(ncl * cluster, n, generator=g)
    xin = x.cuda()
    out = torch.full((ncl * cluster, n), float("nan"), device="cuda")
    cabi.test_cluster_reduce(xin.data_ptr(), out.data_ptr(), n, cluster, ncl, 0, repeats)
    torch.cuda.synchronize()
    scale = sum(range(1, repeats + 1))
    want = x.view(ncl, cluster, n).sum(1, keepdim=True).expand(ncl, cluster, n).reshape(-1, n) * scale
    assert torch.allclose(out.cpu(), want, rtol=1e-5, atol=1e-5)
    # rank-ordered fold: every CTA of a cluster holds the bit-identical result
    o = out.view(ncl, cluster, n)
    assert torch.equal(o, o[:, :1].expand_as(o))


@pytest.mark.parametrize("cluster", [2, 4, 8])
def test_cluster_reduce_gather(cluster):
    from clusterfusion_b200 import cabi
    n, ncl = 128, 2
    x = torch.randn(ncl * cluster, n)
    out = torch.zeros(ncl * cluster, n * cluster, device="cuda")
    cabi.test_cluster_reduce(x.cuda().data_ptr(), out.data_ptr(), n, cluster, ncl, 4, 1)
    torch.cuda.synchronize()
    want = x.view(ncl, 1, cluster * n).expand(ncl, cluster, cluster * n).reshape(-1, cluster * n)
    assert torch.equal(out.cpu(), want)


@pytest.mark.parametrize("cluster", [1, 2, 4, 8])
def test_cluster_reduce_attn_merge(cluster):
    from clusterfusion_b200 import cabi
    D, ncl = 128, 2
    n = D + 4
    g = torch.Generator().manual_seed(cluster)
    st = torch.zeros(ncl * cluster, n)
    st[:, 0] = torch.randn(ncl * cluster, generator=g) * 4           # m (log2 domain)
    st[:, 1] = torch.rand(ncl * cluster, generator=g) + 0.5          # l
    st[:, 4:] = torch.randn(ncl * cluster, D, generator=g)           # o (unnormalised)
    if cluster > 1:
        st[1, 0] = float("-inf"); st[1, 1] = 0; st[1, 4:] = 0        # an empty partial state
    out = torch.zeros(ncl * cluster, n, device="cuda")
    cabi.test_cluster_reduce(st.cuda().data_ptr(), out.data_ptr(), n, cluster, ncl, 1, 2)
    torch.cuda.synchronize()
    s = st.view(ncl, cluster, n).double()
    M = s[:, :, 0].max(1, keepdim=True).values
    w = torch.exp2(s[:, :, 0] - M)
    L = (s[:, :, 1] * w).sum(1)
    Ov = (s[:, :, 4:] * w[:, :, None]).sum(1)
    got = out.cpu().view(ncl, cluster, n).double()
    for r in range(cluster):
        assert torch.allclose(got[:, r, 0], M[:, 0], rtol=1e-6)
        assert torch.allclose(got[:, r, 1], L, rtol=1e-4, atol=1e-5)
        assert torch.allclose(got[:, r, 4:], Ov, rtol=1e-4, atol=1e-4)


# ---------------------------------------------------------------------------------------------------
# 8-arg chat operator (pybind) vs oracle
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kv_len", [0, 1, 31, 37, 128, 256, 1024, 4096])
def test_chat_operator_vs_oracle(kv_len):
    import clusterfusion
    d = O.make_inputs(S7, kv_len, seed=42 + kv_len, layout="chat")
    want_o, want_k, want_v = O.chat_layer(d["x"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                                          d["rms_w"], d["cos"], d["sin"], n_heads=32, eps=1e-6, mode="eager")
    c = cuda(d)
    o, k, v = clusterfusion.llama_decoder_layer(c["x"].view(1, 1, 4096), c["weight_qkv"], c["weight_o"], c["k_cache"],
                                                c["v_cache"], c["rms_w"], c["cos"], c["sin"])
    torch.cuda.synchronize()
    assert o.shape == (1, 4096) and k.shape == (1, 32, 128) and v.shape == (1, 32, 128)
    assert o.dtype == k.dtype == v.dtype == torch.float16
    assert close(v, want_v)
    assert close_k(k, want_k, pairing="gptj")
    assert close(o, want_o)


@pytest.mark.parametrize("kv_len", [0, 1, 37, 256, 1024, 4096, 16384])
def test_sglang_cabi_vs_oracle(kv_len):
    import cabi_torch as ct
    d = O.make_inputs(S7, kv_len, seed=7 + kv_len, layout="sglang")
    eps = 1e-5
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                          d["rms_w"], eps, d["cos"], d["sin"], n_heads=32, mode="eager")
    c = cuda(d)
    o, r, k, v = ct.sglang(c["x"], c["residual"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"],
                           c["rms_w"], eps, c["cos"], c["sin"], n_heads=32)
    torch.cuda.synchronize()
    assert torch.equal(r.cpu(), want[1])
    assert close(v, want[3])
    assert close_k(k, want[2])
    assert close(o, want[0])
    assert torch.equal(c["residual"].cpu(), d["residual"])     # out-of-place form leaves the input alone


def test_sglang_operator_updates_residual_in_place():
    import clusterfusion
    d = O.make_inputs(S7, 300, seed=3, layout="sglang")
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                          d["rms_w"], 1e-6, d["cos"], d["sin"], n_heads=32, mode="eager")
    c = cuda(d)
    res = c["residual"].clone()
    for _ in range(3):      # exercise scratch / counter reset between launches
        res.copy_(c["residual"])
        o, r, k, v = clusterfusion.llama_decoder_layer_sglang(c["x"], res, c["weight_qkv"], c["weight_o"],
                                                              c["k_cache"], c["v_cache"], c["rms_w"], 1e-6,
                                                              c["cos"], c["sin"])
        torch.cuda.synchronize()
        assert r.data_ptr() == res.data_ptr()
        assert torch.equal(res.cpu(), want[1])
        assert close(o, want[0]) and close_k(k, want[2]) and close(v, want[3])


@pytest.mark.parametrize("name", sorted(p.name for p in GOLDEN.glob("sglang_*w0.02.npz")))
def test_sglang_operator_vs_reference_golden(name):
    """CUDA kernel against the fixture the reference's own pure-torch reference() produced."""
    import clusterfusion
    z = np.load(GOLDEN / name)
    d = O.make_inputs(S7, int(z["kv_len"]), seed=int(z["seed"]), w_scale=float(z["w_scale"]), layout="sglang")
    assert inputs_digest(d) == str(z["digest"])
    c = cuda(d)
    o, r, k, v = clusterfusion.llama_decoder_layer_sglang(c["x"], c["residual"], c["weight_qkv"], c["weight_o"],
                                                          c["k_cache"], c["v_cache"], c["rms_w"], float(z["eps"]),
                                                          c["cos"], c["sin"])
    torch.cuda.synchronize()
    # reference() keeps everything in fp32 (no fp16 rounding of q/k/v); the reference's own bar for this
    # comparison is output < 5e-2, residual < 1e-3, k/v < 1e-2 (tests/test_llama_tilelang.py:100).
    # We hold the north-star tolerance wherever there are enough keys to average the q/k/v rounding.
    tol = 1e-3 if int(z["kv_len"]) >= 37 else 2.5e-3
    assert close(o, torch.from_numpy(z["out"]), rtol=tol, atol=tol)
    assert torch.equal(r.cpu(), torch.from_numpy(z["residual_out"]))
    # k/v: the kernel rounds the normalised input to fp16 where the eager fp16 model does (that is the
    # flavour the north-star names and test_sglang_cabi_vs_oracle holds to 1e-3); reference() never rounds,
    # which moves single k/v elements by up to ~2e-3.  Still 3x tighter than the reference's own 1e-2.
    assert close(k, torch.from_numpy(z["k"]), rtol=2e-3, atol=6e-3)
    assert close(v, torch.from_numpy(z["v"]), rtol=2e-3, atol=3e-3)


@pytest.mark.parametrize("name", sorted(p.name for p in GOLDEN.glob("chat_fp*_kv*.npz") if "gqa" not in p.name and "70b" not in p.name))
def test_chat_operator_vs_reference_golden(name):
    """CUDA kernel against what the reference's eager Attention module produced on CPU."""
    import clusterfusion
    z = np.load(GOLDEN / name)
    kv = int(z["kv_len"])
    d = O.make_inputs(S7, kv, seed=int(z["seed"]), w_scale=float(z["w_scale"]), layout="sglang")
    assert inputs_digest(d) == str(z["digest"])
    wq, wk, wv = d["weight_qkv"].split([4096, 4096, 4096], 0)
    wqkv_T = torch.cat([wq.t(), wk.t(), wv.t()], 0).contiguous().cuda()
    wo_T = d["weight_o"].t().contiguous().cuda()
    cos, sin = torch.from_numpy(z["cos"]).cuda(), torch.from_numpy(z["sin"]).cuda()
    o, k, v = clusterfusion.llama_decoder_layer(d["x"].cuda(), wqkv_T, wo_T, d["k_cache"].cuda(), d["v_cache"].cuda(),
                                                d["rms_w"].cuda(), cos, sin)
    torch.cuda.synchronize()
    tol = 1e-3 if kv >= 37 else 2.5e-3
    assert close(o, torch.from_numpy(z["out"]), rtol=tol, atol=tol)
    if str(z["dtype"]) == "float16":     # the reference ran natively in fp16: same rounding points as the kernel
        assert close_k(k, torch.from_numpy(z["k"]))
        assert close(v, torch.from_numpy(z["v"]))
    else:                                # fp32 run of the reference: see the note in the sglang golden test
        assert close(k, torch.from_numpy(z["k"]), rtol=2e-3, atol=6e-3)
        assert close(v, torch.from_numpy(z["v"]), rtol=2e-3, atol=3e-3)


# ---------------------------------------------------------------------------------------------------
# 15-arg paged batch operator
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("lens", [[5], [0, 33, 1], [700, 64, 0, 129]])
def test_paged_operator_vs_oracle(lens):
    import clusterfusion
    bs = len(lens)
    nslots = sum(lens) + bs + 17
    d = O.make_inputs(S7, nslots, seed=100 + bs, layout="sglang", bs=bs)
    g = torch.Generator().manual_seed(1)
    slots = torch.randperm(nslots, generator=g)
    indptr, indices, off = [0], [], 0
    for L in lens:
        indices += slots[off:off + L + 1].tolist()
        off += L + 1
        indptr.append(len(indices))
    indptr = torch.tensor(indptr, dtype=torch.int32)
    indices = torch.tensor(indices, dtype=torch.int32)
    positions = torch.tensor(lens, dtype=torch.int64)
    maxpos = max(lens) + 1
    cos_sin = torch.stack([torch.cat([O.rope_angles(p).cos(), O.rope_angles(p).sin()]) for p in range(maxpos)])
    kp, vp = d["k_cache"].clone(), d["v_cache"].clone()
    want_o, want_r = O.paged_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], indptr, indices, kp, vp,
                                   d["rms_w"], 1e-5, positions, cos_sin, n_heads=32, mode="eager")
    c = cuda(d)
    n_layers, layer_id = 3, 1
    kpools = [torch.zeros_like(c["k_cache"]) for _ in range(n_layers)]
    vpools = [torch.zeros_like(c["v_cache"]) for _ in range(n_layers)]
    kpools[layer_id].copy_(c["k_cache"]); vpools[layer_id].copy_(c["v_cache"])
    kptrs = torch.tensor([t.data_ptr() for t in kpools], dtype=torch.uint64).cuda()
    vptrs = torch.tensor([t.data_ptr() for t in vpools], dtype=torch.uint64).cuda()
    out = torch.full((bs, 4096), float("nan"), dtype=torch.float16, device="cuda")
    rout = torch.full((bs, 4096), float("nan"), dtype=torch.float16, device="cuda")
    clusterfusion.llama_decoder_layer_batch_decode_sglang(
        out, rout, c["x"], c["residual"], c["weight_qkv"], c["weight_o"], indptr.cuda(), indices.cuda(),
        kptrs, vptrs, layer_id, c["rms_w"], 1e-5, positions.cuda(), cos_sin.cuda())
    torch.cuda.synchronize()
    assert torch.equal(rout.cpu(), want_r)
    assert close(out, want_o)
    assert close_k(kpools[layer_id], kp)                # appended K rows (post-RoPE), everything else untouched
    assert close(vpools[layer_id], vp)
    touched = {int(indices[indptr[b + 1] - 1]) for b in range(bs)}
    keep = [i for i in range(nslots) if i not in touched]
    assert torch.equal(kpools[layer_id].cpu()[keep], d["k_cache"][keep])
    assert torch.equal(kpools[0].cpu(), torch.zeros_like(d["k_cache"]))
    # the README spelling of the same call (reference README.md:55-75)
    out2 = torch.empty_like(out); rout2 = torch.empty_like(rout)
    kpools[layer_id].copy_(c["k_cache"]); vpools[layer_id].copy_(c["v_cache"])
    clusterfusion.llama_decoder_layer(out2, rout2, c["x"], c["residual"], c["weight_qkv"], c["weight_o"],
                                      indptr.cuda(), indices.cuda(), kptrs, vptrs, layer_id, c["rms_w"], 1e-5,
                                      positions.cuda(), cos_sin.cuda())
    torch.cuda.synchronize()
    assert close(out2, want_o)


@pytest.mark.parametrize("lens", [[40, 0, 513, 7, 128, 1], [300] * 8, [17, 2], [5000, 17, 3000, 1, 2049],
                                  [64, 0, 1, 900, 33, 16, 15, 17, 700, 2, 31]])
def test_batched_paged_kernel_matches_per_request_kernel_and_oracle(lens):
    """Row f3: the batched kernels (weights streamed once per chunk of 4 requests up to batch 4, of 8 from batch 5; ragged tail) against
    the oracle and against the per-request launch (CF_FLAG_PER_REQUEST, the reference's grid shape)."""
    import cabi_torch as ct
    from clusterfusion_b200 import cabi
    bs = len(lens)
    nslots = sum(lens) + bs + 5
    d = O.make_inputs(S7, nslots, seed=300 + bs, layout="sglang", bs=bs)
    slots = torch.randperm(nslots, generator=torch.Generator().manual_seed(4))
    indptr, indices, off = [0], [], 0
    for L in lens:
        indices += slots[off:off + L + 1].tolist(); off += L + 1; indptr.append(len(indices))
    indptr = torch.tensor(indptr, dtype=torch.int32); indices = torch.tensor(indices, dtype=torch.int32)
    positions = torch.tensor(lens, dtype=torch.int64)
    cos_sin = torch.stack([torch.cat([O.rope_angles(p).cos(), O.rope_angles(p).sin()]) for p in range(max(lens) + 1)])
    kp, vp = d["k_cache"].clone(), d["v_cache"].clone()
    want_o, want_r = O.paged_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], indptr, indices, kp, vp,
                                   d["rms_w"], 1e-5, positions, cos_sin, n_heads=32, mode="eager")
    c = cuda(d)
    res = {}
    for name, flags in (("batched", 0), ("per_request", cabi.CF_FLAG_PER_REQUEST)):
        kpool, vpool = c["k_cache"].clone(), c["v_cache"].clone()
        kptrs = torch.tensor([kpool.data_ptr()], dtype=torch.uint64).cuda()
        vptrs = torch.tensor([vpool.data_ptr()], dtype=torch.uint64).cuda()
        out = torch.full((bs, 4096), float("nan"), dtype=torch.float16, device="cuda"); rout = torch.empty_like(out)
        for _ in range(3):        # repeated launches on one workspace: scratch / counters must come back zeroed
            kpool.copy_(c["k_cache"]); vpool.copy_(c["v_cache"])
            ct.paged(out, rout, c["x"], c["residual"], c["weight_qkv"], c["weight_o"], indptr.cuda(), indices.cuda(), kptrs, vptrs,
                     0, c["rms_w"], 1e-5, positions.cuda(), cos_sin.cuda(), n_heads=32, flags=flags)
        torch.cuda.synchronize()
        assert torch.equal(rout.cpu(), want_r), name
        assert close(out, want_o), name
        assert close_k(kpool, kp) and close(vpool, vp), name
        res[name] = out
    assert close(res["batched"], res["per_request"])


# ---------------------------------------------------------------------------------------------------
# grouped-query attention: Llama-3-8B (32 Q / 8 KV) and the Llama-2-70B head-parallel shards
# ---------------------------------------------------------------------------------------------------
S8 = O.LayerShape(4096, 32, 8)
S70 = O.LayerShape(8192, 64, 8)


@pytest.mark.parametrize("kv_len", [0, 1, 100, 1000, 8192])
def test_gqa_llama3_8b_operator_vs_oracle(kv_len):
    import clusterfusion
    d = O.make_inputs(S8, kv_len, seed=80 + kv_len, layout="sglang", theta=500000.0)
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                          d["rms_w"], 1e-5, d["cos"], d["sin"], n_heads=32, n_kv_heads=8, mode="eager")
    c = cuda(d)
    res = c["residual"].clone()
    o, r, k, v = clusterfusion.llama_decoder_layer_sglang(c["x"], res, c["weight_qkv"], c["weight_o"], c["k_cache"],
                                                          c["v_cache"], c["rms_w"], 1e-5, c["cos"], c["sin"])
    torch.cuda.synchronize()
    assert k.shape == (1, 8, 128) and v.shape == (1, 8, 128)
    assert torch.equal(r.cpu(), want[1])
    assert close(v, want[3])
    assert close_k(k, want[2])
    assert close(o, want[0])


@pytest.mark.parametrize("shape,kv_len", [(S8, 8192), (S8, 17), (S70, 64), (O.LayerShape(8192, 16, 2), 3000),
                                          (O.LayerShape(8192, 8, 1), 1000), (O.LayerShape(4096, 8, 2), 555)])
def test_gqa_group_kernel_shapes_vs_oracle(shape, kv_len):
    """The grouped-query kernel (G CTAs per group with L2 exchanges, G = 8 / 16 / 32 / 64 depending on the shape) against the
    oracle: Llama-3-8B, the full Llama-2-70B layer and its 1/4 and 1/8 head shards, and a small 8/2 shape."""
    import cabi_torch as ct
    from clusterfusion_b200 import cabi
    d = O.make_inputs(shape, kv_len, seed=kv_len + shape.n_heads, layout="sglang")
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                          d["rms_w"], 1e-5, d["cos"], d["sin"], n_heads=shape.n_heads, n_kv_heads=shape.n_kv_heads, mode="eager")
    c = cuda(d)
    for flags in (0,):
        o, r, k, v = ct.sglang(c["x"], c["residual"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"],
                               1e-5, c["cos"], c["sin"], n_heads=shape.n_heads, n_kv_heads=shape.n_kv_heads, flags=flags)
        torch.cuda.synchronize()
        assert torch.equal(r.cpu(), want[1])
        assert close(v, want[3])
        assert close_k(k, want[2])
        assert close(o, want[0])


def test_gqa_group_kernel_long_context_32k():
    """Llama-3-8B shapes at kv 32K: 128 K/V tiles per CTA through the tensor-core attention loop (many online-softmax
    rescales per warp), against the oracle run in full."""
    import cabi_torch as ct
    kv = 32768
    d = O.make_inputs(S8, kv, seed=32, layout="sglang", theta=500000.0)
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                          d["rms_w"], 1e-5, d["cos"], d["sin"], n_heads=32, n_kv_heads=8, mode="eager")
    c = cuda(d)
    o, r, k, v = ct.sglang(c["x"], c["residual"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"],
                           1e-5, c["cos"], c["sin"], n_heads=32, n_kv_heads=8)
    torch.cuda.synchronize()
    assert close(o, want[0]) and close(v, want[3]) and close_k(k, want[2])


def test_gqa_attention_is_a_convex_combination():
    """Size-independent property of the tensor-core loop: if every cached V row of a KV head is the same vector u and the
    new token's V row equals u as well, the attention output of all 4 query heads of that KV head is exactly u whatever
    the scores are (the fp16-rounded probabilities are normalised by their own sum).  Checked through Wo = identity."""
    import cabi_torch as ct
    kv, H = 5000, 4096
    g = torch.Generator().manual_seed(11)
    x = torch.randn(1, H, generator=g).half(); res = torch.zeros(1, H).half()
    wq = (torch.randn(32 * 128, H, generator=g) * 0.05).half()
    wk = torch.zeros(8 * 128, H).half()                        # new token: k = 0 -> score 0, negligible against 5000 keys
    wv = torch.zeros(8 * 128, H).half()                        #            v = 0
    rms = torch.ones(H).half()
    kc = torch.randn(kv, 8 * 128, generator=g).half()
    cvec = torch.randn(8 * 128, generator=g).half()
    vc = cvec.expand(kv, 8 * 128).contiguous()
    wo = torch.eye(H).half()                                   # hidden == Hq * 128: out = attention output
    ang = O.rope_angles(kv, theta=500000.0)
    o, r, k, v = ct.sglang(x.cuda(), res.cuda(), torch.cat([wq, wk, wv], 0).contiguous().cuda(), wo.cuda(), kc.cuda(), vc.cuda(),
                           rms.cuda(), 1e-5, ang.cos().cuda(), ang.sin().cuda(), n_heads=32, n_kv_heads=8)
    torch.cuda.synchronize()
    # out[head h] = (1 - p_new) * cvec[kv head] + p_new * 0; p_new = 1 / (1 + sum_s exp(score_s)) < 1e-3 here
    want = cvec.view(8, 1, 128).expand(8, 4, 128).reshape(1, H).float()
    got = o.float().cpu()
    sel = want.abs() > 0.25
    assert sel.sum() > 1000
    # every selected element is shrunk by the same factor (1 - p_new) of its head, within fp16 rounding
    per_head = (got[sel] / want[sel])
    assert float(per_head.min()) > 0.995 and float(per_head.max()) < 1.002


def test_gqa_group_kernel_repeatability_and_workspace_reset():
    """200 back-to-back launches on one workspace (in-place residual included): the L2 exchange buffers and group counters
    must come back zeroed every time, and results must agree to 1 fp16 ulp (fp32 cross-group red order is the only
    order-dependent step)."""
    import cabi_torch as ct
    d = cuda(O.make_inputs(S8, 2048, seed=9, layout="sglang", theta=500000.0))
    outs = []
    for i in range(200):
        res = d["residual"].clone() if i % 2 else d["residual"]
        ro = res if i % 2 else None                       # odd launches update the residual in place
        o, r, k, v = ct.sglang(d["x"], res, d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], 1e-5,
                               d["cos"], d["sin"], n_heads=32, n_kv_heads=8, residual_out=ro)
        outs.append((o, r))
    torch.cuda.synchronize()
    ref = outs[0][0].float()
    ulp = torch.maximum(ref.abs() * 2 ** -10, torch.full_like(ref, 2 ** -24))
    for o, r in outs[1:]:
        assert bool(((o.float() - ref).abs() <= ulp).all())
        assert torch.equal(r, outs[0][1])
    # workspace: header [epoch, finalised-slice count] | scratch fp32[hidden] | counters u32[32] | exchange words | counters
    ws = ct.workspace(4096, 1, d["x"].device)
    hdr = ws[:8].view(torch.int32).cpu()
    assert int(hdr[0]) >= 200 and int(hdr[1]) == 0          # one epoch per group-kernel launch on this workspace
    assert int(ws[8:12].view(torch.int32).item()) == 0      # header[2]: no exchange poll ever timed out
    assert int(ws[256:256 + 4096 * 4 + 32 * 4].count_nonzero()) == 0
    tail = (4096 // 128) * 4096 * 8                          # batch-1 output words live at the end; counters just before
    assert int(ws[-tail - 128 * 4:-tail].count_nonzero()) == 0


def _gptj_to_neox_perm():
    perm = torch.empty(128, dtype=torch.long)         # neox index j holds gptj index perm[j]
    perm[:64] = torch.arange(0, 128, 2)
    perm[64:] = torch.arange(1, 128, 2)
    return perm


@pytest.mark.parametrize("name,shape", [("chat_fp32_gqa8_kv100.npz", S8)])
def test_gqa_operator_vs_reference_eager_golden(name, shape):
    """The GQA fixture came from the reference's eager Attention (GPT-J pair RoPE).  q.k is invariant under a common
    permutation of the head dimension, so permuting Wq / Wk rows and the K cache columns from pair order to
    rotate-half order lets the NeoX kernel reproduce the same layer output."""
    import cabi_torch as ct
    z = np.load(GOLDEN / name)
    kv = int(z["kv_len"])
    d = O.make_inputs(shape, kv, seed=int(z["seed"]), w_scale=float(z["w_scale"]), layout="sglang")
    assert inputs_digest(d) == str(z["digest"])
    perm = _gptj_to_neox_perm()
    wq, wk, wv = d["weight_qkv"].split([shape.q_dim, shape.kv_dim, shape.kv_dim], 0)
    wq_p = wq.view(shape.n_heads, 128, -1)[:, perm].reshape(shape.q_dim, -1)
    wk_p = wk.view(shape.n_kv_heads, 128, -1)[:, perm].reshape(shape.kv_dim, -1)
    kc_p = d["k_cache"].view(kv, shape.n_kv_heads, 128)[:, :, perm].reshape(kv, -1)
    ang = O.rope_angles(kv)
    zero = torch.zeros_like(d["x"])
    o, r, k, v = ct.sglang(d["x"].cuda(), zero.cuda(), torch.cat([wq_p, wk_p, wv], 0).contiguous().cuda(),
                           d["weight_o"].cuda(), kc_p.contiguous().cuda(), d["v_cache"].cuda(), d["rms_w"].cuda(),
                           float(z["eps"]), ang.cos().cuda(), ang.sin().cuda(), n_heads=shape.n_heads,
                           n_kv_heads=shape.n_kv_heads)
    torch.cuda.synchronize()
    assert close(o, torch.from_numpy(z["out"]))
    k_ref = torch.from_numpy(z["k"]).view(shape.n_kv_heads, 128)[:, perm]
    assert close(k.view(shape.n_kv_heads, 128), k_ref, rtol=2e-3, atol=6e-3)
    assert close(v, torch.from_numpy(z["v"]), rtol=2e-3, atol=3e-3)


def test_gqa_paged_operator_vs_oracle():
    import clusterfusion
    lens = [300, 0, 45]
    bs = len(lens)
    nslots = sum(lens) + bs + 9
    d = O.make_inputs(S8, nslots, seed=123, layout="sglang", bs=bs)
    slots = torch.randperm(nslots, generator=torch.Generator().manual_seed(2))
    indptr, indices, off = [0], [], 0
    for L in lens:
        indices += slots[off:off + L + 1].tolist(); off += L + 1; indptr.append(len(indices))
    indptr = torch.tensor(indptr, dtype=torch.int32); indices = torch.tensor(indices, dtype=torch.int32)
    positions = torch.tensor(lens, dtype=torch.int64)
    cos_sin = torch.stack([torch.cat([O.rope_angles(p, theta=500000.0).cos(), O.rope_angles(p, theta=500000.0).sin()])
                           for p in range(max(lens) + 1)])
    kp, vp = d["k_cache"].clone(), d["v_cache"].clone()
    want_o, want_r = O.paged_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], indptr, indices, kp, vp,
                                   d["rms_w"], 1e-5, positions, cos_sin, n_heads=32, n_kv_heads=8, mode="eager")
    c = cuda(d)
    kpool, vpool = c["k_cache"].clone(), c["v_cache"].clone()
    kptrs = torch.tensor([kpool.data_ptr()], dtype=torch.uint64).cuda()
    vptrs = torch.tensor([vpool.data_ptr()], dtype=torch.uint64).cuda()
    out = torch.empty(bs, 4096, dtype=torch.float16, device="cuda"); rout = torch.empty_like(out)
    clusterfusion.llama_decoder_layer_batch_decode_sglang(out, rout, c["x"], c["residual"], c["weight_qkv"], c["weight_o"],
                                                          indptr.cuda(), indices.cuda(), kptrs, vptrs, 0, c["rms_w"], 1e-5,
                                                          positions.cuda(), cos_sin.cuda())
    torch.cuda.synchronize()
    assert torch.equal(rout.cpu(), want_r)
    assert close(out, want_o)
    assert close_k(kpool, kp) and close(vpool, vp)


@pytest.mark.parametrize("world,kv", [(2, 777), (4, 777), (8, 777), (2, 1024), (8, 1024), (4, 16384), (8, 16384)])
def test_llama2_70b_head_parallel_shards_sum_to_full_layer(world, kv):
    """Every rank's kernel emits its fp32 O-projection partial (CF_FLAG_OUT_FP32_PARTIAL); the sum over ranks
    (what the one all-reduce per layer computes) must equal the full 64/8-head layer of the oracle -- at a ragged kv and at
    the two kv lengths bench.py times the sharded layer at (1024, 16384).  Ranks emulated in turn on one GPU, so this runs on
    the driver's 1-GPU box; the real 2-GPU exchange is tests/test_gpu_multi.py and bench.py's parity_ok."""
    from clusterfusion_b200 import cabi
    from clusterfusion_b200 import sharded
    import cabi_torch as ct
    d = O.make_inputs(S70, kv, seed=70, layout="sglang")
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                          d["rms_w"], 1e-5, d["cos"], d["sin"], n_heads=64, n_kv_heads=8, mode="eager")
    c = cuda(d)
    total = torch.zeros(1, 8192, dtype=torch.float32, device="cuda")
    k_all, v_all = [], []
    ws = ct.workspace(8192, 1, c["x"].device)
    for rank in range(world):
        sh = sharded.shard_layer(c["weight_qkv"], c["weight_o"], 64, 8, rank, world)
        kc = sharded.shard_kv(c["k_cache"], 8, rank, world); vc = sharded.shard_kv(c["v_cache"], 8, rank, world)
        part = torch.empty(1, 8192, dtype=torch.float32, device="cuda")
        ro = torch.empty(1, 8192, dtype=torch.float16, device="cuda")
        nq, nkv = 64 // world, 8 // world
        kn = torch.empty(nkv * 128, dtype=torch.float16, device="cuda"); vn = torch.empty_like(kn)
        a = cabi.CfLlamaArgs(variant=1, flags=cabi.CF_FLAG_OUT_FP32_PARTIAL, hidden=8192, n_q_heads=nq, n_kv_heads=nkv,
                             head_dim=128, batch=1, kv_len=kv, eps=1e-5, x=c["x"].data_ptr(), residual_in=c["residual"].data_ptr(),
                             residual_out=ro.data_ptr(), w_qkv=sh["w_qkv"].data_ptr(), w_o=sh["w_o"].data_ptr(),
                             rms_w=c["rms_w"].data_ptr(), out=part.data_ptr(), k_new=kn.data_ptr(), v_new=vn.data_ptr(),
                             k_cache=kc.data_ptr(), v_cache=vc.data_ptr(), cos=c["cos"].data_ptr(), sin=c["sin"].data_ptr(),
                             workspace=ws.data_ptr())
        cabi.launch(a, ct.stream_handle())
        torch.cuda.synchronize()
        total += part
        k_all.append(kn); v_all.append(vn)
        assert torch.equal(ro.cpu(), want[1])
    assert close(total.half(), want[0])
    assert close_k(torch.cat(k_all), want[2])
    assert close(torch.cat(v_all), want[3])


# ---------------------------------------------------------------------------------------------------
# properties at full BASELINE sizes (no oracle needed) + repeatability
# ---------------------------------------------------------------------------------------------------
def test_repeatability_and_workspace_reset():
    """The reference watches 10 000 runs for atomic-order noise (tests/test_llama.py:22).  Here the only
    order-dependent step is the fp32 cross-head red: repeats must agree to within 1 fp16 ulp."""
    import clusterfusion
    d = cuda(O.make_inputs(S7, 1024, seed=1, layout="chat"))
    outs = []
    for _ in range(300):
        o, k, v = clusterfusion.llama_decoder_layer(d["x"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                                                    d["rms_w"], d["cos"], d["sin"])
        outs.append(o)
    torch.cuda.synchronize()
    ref = outs[0].float()
    ulp = torch.maximum(ref.abs() * 2 ** -10, torch.full_like(ref, 2 ** -24))
    for o in outs[1:]:
        assert bool(((o.float() - ref).abs() <= ulp).all())
    assert bool(torch.equal(k, k)) and not torch.isnan(outs[-1]).any()


def test_ll_out_flag_is_bitwise_reproducible():
    """CF_FLAG_LL_OUT: the cross-head O reduction goes through flag-in-data words summed in head order -> 100 launches
    are bit-identical (the default red path is only reproducible to 1 fp16 ulp) and match the oracle."""
    import cabi_torch as ct
    from clusterfusion_b200 import cabi
    d = O.make_inputs(S7, 777, seed=3, layout="chat")
    want = O.chat_layer(d["x"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], d["cos"], d["sin"],
                        n_heads=32, eps=1e-6, mode="eager")
    c = cuda(d)
    outs = [ct.chat(c["x"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], c["cos"], c["sin"],
                    flags=cabi.CF_FLAG_LL_OUT)[0] for _ in range(100)]
    torch.cuda.synchronize()
    assert close(outs[0], want[0])
    assert all(torch.equal(o, outs[0]) for o in outs[1:])


def test_softmax_shift_invariance_16k():
    """Size-independent property at the headline size: attention over a cache whose V rows are all the same
    vector u must return exactly Wo.u regardless of K (softmax weights sum to 1)."""
    import clusterfusion
    kv = 16384
    d = O.make_inputs(S7, 8, seed=5, layout="chat")
    c = cuda(d)
    u = torch.randn(1, 4096, generator=torch.Generator().manual_seed(0)).half().cuda()
    kc = torch.randn(kv, 4096, device="cuda", dtype=torch.float16)
    vc = u.expand(kv, 4096).contiguous()
    # make the new token negligible: its score is finite, but 16K keys dominate only if it is not huge; use
    # the general identity instead: o = Wo^T-proj of (sum_s p_s u + p_new v_new)
    o, k, v = clusterfusion.llama_decoder_layer(c["x"], c["weight_qkv"], c["weight_o"], kc, vc, c["rms_w"], c["cos"], c["sin"])
    # oracle on the collapsed problem: all 16K cache rows share V=u, so the result equals attention over the
    # true K with that V; compare with the oracle run on a strided subsample is not exact -> run the oracle in full
    want_o, _, _ = O.chat_layer(d["x"], d["weight_qkv"], d["weight_o"], kc.cpu(), vc.cpu(), d["rms_w"], d["cos"], d["sin"],
                                n_heads=32, eps=1e-6, mode="eager")
    torch.cuda.synchronize()
    assert close(o, want_o)


def test_pdl_chain_matches_plain_chain():
    """CF_FLAG_PDL lets layer l+1 start streaming its weights while layer l finishes; results must not change.
    The chain is the real decoder data flow (sglang form: x_{l+1} = o_l, residual_{l+1} = x_l + residual_l), so every
    launch consumes what the previous launch produced -- reading x before the previous kernel completed would show."""
    from clusterfusion_b200 import cabi
    import cabi_torch as ct
    nl, kv = 6, 700
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device="cuda") * sc).half()
    layers = [dict(w_qkv=r(3 * 4096, 4096, sc=0.02), w_o=r(4096, 4096, sc=0.02), k=r(kv, 4096), v=r(kv, 4096),
                   rms=(1 + 0.1 * r(4096).float()).half()) for _ in range(nl)]
    x0, res0 = r(1, 4096), r(1, 4096)
    cos = torch.rand(64, device="cuda"); sin = torch.rand(64, device="cuda")
    ws = ct.workspace(4096, 1, x0.device)

    def chain(flags):
        outs, h, res = [], x0, res0
        for lay in layers:
            o = torch.empty(1, 4096, dtype=torch.float16, device="cuda")
            ro = torch.empty(1, 4096, dtype=torch.float16, device="cuda")
            kn = torch.empty(4096, dtype=torch.float16, device="cuda"); vn = torch.empty_like(kn)
            a = cabi.CfLlamaArgs(variant=1, flags=flags, hidden=4096, n_q_heads=32, n_kv_heads=32, head_dim=128, batch=1,
                                 kv_len=kv, eps=1e-5, x=h.data_ptr(), residual_in=res.data_ptr(), residual_out=ro.data_ptr(),
                                 w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(), rms_w=lay["rms"].data_ptr(),
                                 out=o.data_ptr(), k_new=kn.data_ptr(), v_new=vn.data_ptr(), k_cache=lay["k"].data_ptr(),
                                 v_cache=lay["v"].data_ptr(), cos=cos.data_ptr(), sin=sin.data_ptr(), workspace=ws.data_ptr())
            cabi.launch(a, ct.stream_handle())
            outs += [o, ro, kn, vn]
            h, res = o, ro
        torch.cuda.synchronize()
        return outs

    plain = chain(0)
    side = torch.cuda.Stream()
    for it in range(6):
        if it < 3:
            pdl = chain(cabi.CF_FLAG_PDL)
        else:
            with torch.cuda.stream(side):
                pdl = chain(cabi.CF_FLAG_PDL)
        for a, b in zip(plain, pdl):
            assert close(b, a, rtol=2e-3, atol=4e-3)


# ---------------------------------------------------------------------------------------------------
# fused FFN half-layer
# ---------------------------------------------------------------------------------------------------
def _ffn_inputs(hidden, ffn, seed):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *sh, sc=1.0: (torch.randn(*sh, generator=g, dtype=torch.float32) * sc).half()
    return dict(x=rn(1, hidden), residual=rn(1, hidden), w13=rn(2 * ffn, hidden, sc=0.02), w2t=rn(ffn, hidden, sc=0.02),
                rms=(1.0 + 0.1 * torch.randn(hidden, generator=g)).half())


@pytest.mark.parametrize("hidden,ffn", [(4096, 11008), (4096, 14336), (8192, 3584), (1024, 16), (2048, 4800)])
def test_ffn_operator_vs_oracle(hidden, ffn):
    import clusterfusion
    d = _ffn_inputs(hidden, ffn, seed=ffn)
    want_o, want_r = O.ffn_layer(d["x"], d["residual"], d["w13"], d["w2t"], d["rms"], 1e-5, mode="eager")
    c = cuda(d)
    for _ in range(3):                                  # workspace must come back zeroed every time
        o, r = clusterfusion.llama_ffn_layer(c["x"], c["residual"], c["w13"], c["w2t"], c["rms"], 1e-5)
        torch.cuda.synchronize()
        assert torch.equal(r.cpu(), want_r)
        assert close(o, want_o, rtol=2e-3, atol=2e-3)   # |out| ~ 2: 1 fp16 ulp = 2e-3 (the attention bar of 1e-3 is for |out| < 1)
    # in-place residual stream
    res = c["residual"].clone()
    o2 = torch.empty_like(o)
    clusterfusion.llama_ffn_layer_out(o2, res, c["x"], res, c["w13"], c["w2t"], c["rms"], 1e-5)
    torch.cuda.synchronize()
    assert torch.equal(res.cpu(), want_r) and close(o2, want_o, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("name", sorted(p.name for p in GOLDEN.glob("ffn_*.npz")))
def test_ffn_operator_vs_reference_feedforward_golden(name):
    """CUDA FFN kernel against the reference's FeedForward + RMSNorm modules run on CPU (model.py:407-448)."""
    import clusterfusion
    from oracle.gen_golden import ffn_inputs
    z = np.load(GOLDEN / name)
    d = ffn_inputs(int(z["seed"]))
    assert inputs_digest(d) == str(z["digest"])
    w13 = torch.cat([d["w1"], d["w3"]], 0).contiguous().cuda()
    w2t = d["w2"].t().contiguous().cuda()
    o, r = clusterfusion.llama_ffn_layer(d["x"].cuda(), d["residual"].cuda(), w13, w2t, d["rms"].cuda(), float(z["eps"]))
    torch.cuda.synchronize()
    tol = 2e-3 if str(z["dtype"]) == "float16" else 4e-3    # fp32 run of the reference never rounds the activations
    assert close(o, torch.from_numpy(z["out"]), rtol=tol, atol=tol)


def test_attention_and_ffn_share_one_workspace_in_a_chain():
    """attention op -> FFN op -> attention op ... on one stream, PDL on, one workspace: a full decoder stack."""
    from clusterfusion_b200 import cabi
    import cabi_torch as ct
    nl, kv, H, F = 4, 200, 4096, 11008
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device="cuda") * sc).half()
    L = [dict(w_qkv=r(3 * H, H, sc=0.02), w_o=r(H, H, sc=0.02), k=r(kv, H), v=r(kv, H), n1=(1 + 0.1 * r(H).float()).half(),
              w13=r(2 * F, H, sc=0.02), w2t=r(F, H, sc=0.02), n2=(1 + 0.1 * r(H).float()).half()) for _ in range(nl)]
    x0, res0 = r(1, H), r(1, H)
    cos = torch.rand(64, device="cuda"); sin = torch.rand(64, device="cuda")
    ws = ct.workspace(H, 1, x0.device)

    def chain(flags):
        x, res, keep = x0, res0, []
        for l in L:
            o = torch.empty(1, H, dtype=torch.float16, device="cuda"); ro = torch.empty_like(o)
            kn = torch.empty(H, dtype=torch.float16, device="cuda"); vn = torch.empty_like(kn)
            a = cabi.CfLlamaArgs(variant=1, flags=flags, hidden=H, n_q_heads=32, n_kv_heads=32, head_dim=128, batch=1, kv_len=kv,
                                 eps=1e-5, x=x.data_ptr(), residual_in=res.data_ptr(), residual_out=ro.data_ptr(),
                                 w_qkv=l["w_qkv"].data_ptr(), w_o=l["w_o"].data_ptr(), rms_w=l["n1"].data_ptr(), out=o.data_ptr(),
                                 k_new=kn.data_ptr(), v_new=vn.data_ptr(), k_cache=l["k"].data_ptr(), v_cache=l["v"].data_ptr(),
                                 cos=cos.data_ptr(), sin=sin.data_ptr(), workspace=ws.data_ptr())
            cabi.launch(a, ct.stream_handle())
            f = torch.empty_like(o); rf = torch.empty_like(o)
            b = cabi.CfFfnArgs(flags=flags, hidden=H, ffn=F, eps=1e-5, x=o.data_ptr(), residual_in=ro.data_ptr(),
                               w_gate_up=l["w13"].data_ptr(), w_down_t=l["w2t"].data_ptr(), rms_w=l["n2"].data_ptr(),
                               out=f.data_ptr(), residual_out=rf.data_ptr(), workspace=ws.data_ptr())
            cabi.launch_ffn(b, ct.stream_handle())
            keep += [o, ro, f, rf]
            x, res = f, rf
        torch.cuda.synchronize()
        return keep

    plain = chain(0)
    # oracle for the first layer
    l = L[0]
    wo, wr, _, _ = O.sglang_layer(x0.cpu(), res0.cpu(), l["w_qkv"].cpu(), l["w_o"].cpu(), l["k"].cpu(), l["v"].cpu(), l["n1"].cpu(),
                                  1e-5, cos.cpu(), sin.cpu(), n_heads=32, mode="eager")
    assert close(plain[0], wo) and torch.equal(plain[1].cpu(), wr)
    fo, fr = O.ffn_layer(plain[0].cpu(), plain[1].cpu(), l["w13"].cpu(), l["w2t"].cpu(), l["n2"].cpu(), 1e-5, mode="eager")
    assert close(plain[2], fo, rtol=2e-3, atol=2e-3) and torch.equal(plain[3].cpu(), fr)
    for _ in range(4):
        pdl = chain(cabi.CF_FLAG_PDL)
        for j, (a_, b_) in enumerate(zip(plain, pdl)):
            # fp32 atomic order moves last fp16 bits and the 8 chained kernels (random weights) amplify them layer by
            # layer: tight for the first layer, a growing share of stragglers deeper down.  A real ordering bug (reading
            # x / residual / workspace before the previous kernel completed) corrupts whole vectors at every depth.
            d_ = (a_.float() - b_.float()).abs()
            bad = d_ > 6e-3 + 3e-3 * a_.float().abs()
            layer = j // 4
            assert float(bad.float().mean()) < (2e-3 if layer == 0 else 3e-2) and float(d_.max()) < 6e-2


def test_decode_engine_fused_matches_eager_attention():
    """Whole-model decode loop (clusterfusion_b200/decode.py): fused attention op vs eager PyTorch attention, same
    random weights, CUDA-graphed and not; a few tokens from position 37."""
    from clusterfusion_b200.decode import LlamaDecodeEngine, ModelShape
    shp = ModelShape(n_layers=3, hidden=4096, n_heads=32, n_kv_heads=32, ffn=1024, vocab=2048)
    toks = {}
    for mode, graph in (("eager", False), ("fused", False), ("fused", True), ("fused+ffn", True)):
        eng = LlamaDecodeEngine(shp, max_seq=128, device="cuda", seed=3, attn="eager" if mode == "eager" else "fused",
                                ffn="fused" if mode == "fused+ffn" else "torch")
        eng.set_position(37)
        if graph:
            eng.capture()
        seq, t = [], 5
        for _ in range(6):
            t = eng.step_host(t)
            seq.append(t)
        toks[(mode, graph)] = seq
        assert int(eng.positions[0]) == 43 and int(eng.indptr[1]) == 44
    assert toks[("fused", False)] == toks[("fused", True)]
    # greedy tokens can legitimately diverge after a near-tie; the first ones must agree
    assert toks[("fused", True)][:3] == toks[("eager", False)][:3]
    assert toks[("fused+ffn", True)][:3] == toks[("eager", False)][:3]


# ---------------------------------------------------------------------------------------------------
# standalone cluster RMSNorm op (row f4)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("batch,hidden", [(64, 8192), (1, 4096), (7, 1024), (3, 16384), (5, 48)])
def test_rmsnorm_operator_vs_oracle(batch, hidden):
    import clusterfusion
    g = torch.Generator().manual_seed(batch * 131 + hidden)
    x = torch.randn(batch, hidden, generator=g).half()
    w = torch.randn(hidden, generator=g).half()
    want = O.rmsnorm_op(x, w, 1e-6)
    got = clusterfusion.rmsnorm(x.cuda(), w.cuda())
    torch.cuda.synchronize()
    # one fp16 rounding of an fp32 product on both sides: at most 1 ulp apart (rsqrt.approx vs torch.rsqrt)
    assert torch.allclose(got.float().cpu(), want.float(), rtol=1e-3, atol=1e-3)
    assert float((got.float().cpu() - want.float()).abs().max()) <= float(want.float().abs().max()) * 2 ** -10


def test_errors_are_loud():
    import clusterfusion
    d = cuda(O.make_inputs(S7, 4, seed=1, layout="chat"))
    with pytest.raises(RuntimeError):
        clusterfusion.llama_decoder_layer(d["x"].float(), d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                                          d["rms_w"], d["cos"], d["sin"])
    with pytest.raises(RuntimeError):
        clusterfusion.llama_decoder_layer(d["x"], d["weight_qkv"][:100], d["weight_o"], d["k_cache"], d["v_cache"],
                                          d["rms_w"], d["cos"], d["sin"])
    with pytest.raises(RuntimeError):
        clusterfusion.llama_decoder_layer(d["x"].cpu(), d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                                          d["rms_w"], d["cos"], d["sin"])
    from clusterfusion_b200 import cabi
    a = cabi.CfLlamaArgs(variant=7)
    with pytest.raises(cabi.CfError) as e:
        cabi.launch(a)
    assert e.value.code == -2


# ---------------------------------------------------------------------------------------------------
# long context (BASELINE configs 2-3: kv 16K / 64K) against the oracle run in full
# ---------------------------------------------------------------------------------------------------
def _paged_case(lens, seed, table="random", nslots_extra=37, shape=None):
    """Inputs + oracle result of a 15-argument call.  table: random = a permutation of the pool; sequential = one run of
    consecutive slots per request; runs = runs of 48/16/5/31/64/1/17 slots with gaps and backward jumps."""
    shape = shape or S7
    bs = len(lens)
    nslots = sum(lens) + bs + nslots_extra
    d = O.make_inputs(shape, nslots, seed=seed, layout="sglang", bs=bs)
    need = sum(lens) + bs
    if table == "random":
        slots = torch.randperm(nslots, generator=torch.Generator().manual_seed(seed)).tolist()
    elif table == "sequential":
        slots = list(range(nslots_extra // 2, nslots_extra // 2 + need))
    else:
        slots, s0, free = [], 0, list(range(nslots))
        runs = (48, 16, 5, 31, 64, 1, 17)
        chunks = []
        while s0 < nslots:
            for r in runs:
                chunks.append(free[s0:s0 + r]); s0 += r
        order = torch.randperm(len(chunks), generator=torch.Generator().manual_seed(seed)).tolist()
        for ci in order:
            slots += chunks[ci]
    indptr, indices, off = [0], [], 0
    for L in lens:
        indices += slots[off:off + L + 1]; off += L + 1; indptr.append(len(indices))
    assert len(set(indices)) == len(indices)
    indptr = torch.tensor(indptr, dtype=torch.int32); indices = torch.tensor(indices, dtype=torch.int32)
    positions = torch.tensor(lens, dtype=torch.int64)
    pos_tab = torch.arange(max(lens) + 1, dtype=torch.float32)[:, None] * O.rope_angles(1)[None, :]
    cos_sin = torch.cat([pos_tab.cos(), pos_tab.sin()], 1).contiguous()
    kp, vp = d["k_cache"].clone(), d["v_cache"].clone()
    want_o, want_r = O.paged_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], indptr, indices, kp, vp,
                                   d["rms_w"], 1e-5, positions, cos_sin, n_heads=shape.n_heads, n_kv_heads=shape.n_kv_heads, mode="eager")
    return d, indptr, indices, positions, cos_sin, kp, vp, want_o, want_r


def _run_paged_cabi(d, indptr, indices, positions, cos_sin, flags=0, host_pool_copy="right", shape=None, inplace=False, fp32_out=False):
    """One 15-argument launch through the C ABI.  host_pool_copy: right = pass the host's copy of the pool addresses
    (tiled / gather4 KV fast paths), none = NULL (row-by-row gather), stale = a WRONG copy (the kernel must notice)."""
    import cabi_torch as CT
    from clusterfusion_b200 import cabi
    shape = shape or S7
    H = shape.hidden
    c = cuda(d)
    bs = d["x"].shape[0]
    kpool, vpool = c["k_cache"].clone(), c["v_cache"].clone()
    decoy_k, decoy_v = torch.zeros_like(kpool), torch.zeros_like(vpool)
    kptrs = torch.tensor([kpool.data_ptr()], dtype=torch.uint64).cuda()
    vptrs = torch.tensor([vpool.data_ptr()], dtype=torch.uint64).cuda()
    out = torch.full((bs, H), float("nan"), dtype=torch.float32 if fp32_out else torch.float16, device="cuda")
    rout = c["residual"] if inplace else torch.full((bs, H), float("nan"), dtype=torch.float16, device="cuda")
    dev_t = (indptr.cuda(), indices.cuda(), positions.cuda(), cos_sin.cuda())
    hk = {"right": kpool, "stale": decoy_k, "none": None}[host_pool_copy]
    hv = {"right": vpool, "stale": decoy_v, "none": None}[host_pool_copy]
    a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_PAGED, flags=flags, hidden=H, n_q_heads=shape.n_heads, n_kv_heads=shape.n_kv_heads,
                         head_dim=128, batch=bs, layer_id=0, eps=1e-5, x=c["x"].data_ptr(), residual_in=c["residual"].data_ptr(),
                         residual_out=rout.data_ptr(), w_qkv=c["weight_qkv"].data_ptr(), w_o=c["weight_o"].data_ptr(),
                         rms_w=c["rms_w"].data_ptr(), out=out.data_ptr(), indptr=dev_t[0].data_ptr(), indices=dev_t[1].data_ptr(),
                         k_pool_ptrs=kptrs.data_ptr(), v_pool_ptrs=vptrs.data_ptr(), positions=dev_t[2].data_ptr(),
                         cos=dev_t[3].data_ptr(), k_cache=None if hk is None else hk.data_ptr(),
                         v_cache=None if hv is None else hv.data_ptr(), workspace=CT.workspace(H, bs, c["x"].device).data_ptr())
    cabi.launch(a, CT.stream_handle())
    torch.cuda.synchronize()
    assert torch.equal(decoy_k.cpu(), torch.zeros_like(d["k_cache"]))
    return out, rout, kpool, vpool


@pytest.mark.parametrize("table", ["random", "sequential", "runs"])
@pytest.mark.parametrize("host_pool_copy", ["right", "none", "stale"])
def test_paged_kv_fetch_paths_agree_with_oracle(table, host_pool_copy):
    """The three ways the per-request paged kernel fills a KV stage -- tiled TMA over 16 consecutive slots, tile::gather4 over
    arbitrary slots, row-by-row bulk copies (ragged tail; no / stale host copy of the pool addresses) -- against the oracle, on
    a random table, a sequential one and one made of runs of 48/16/5/31/64/1/17 slots."""
    from clusterfusion_b200 import cabi
    d, indptr, indices, positions, cos_sin, kp, vp, want_o, want_r = _paged_case([1000], seed=321, table=table)
    out, rout, kpool, vpool = _run_paged_cabi(d, indptr, indices, positions, cos_sin, host_pool_copy=host_pool_copy)
    assert torch.equal(rout.cpu(), want_r)
    assert close(out, want_o)
    assert close_k(kpool, kp) and close(vpool, vp)


@pytest.mark.parametrize("table", ["random", "sequential", "runs"])
@pytest.mark.parametrize("host_pool_copy", ["right", "none", "stale"])
@pytest.mark.parametrize("lens", [[1000], [333, 0, 2100]])
def test_gqa_paged_kv_fetch_paths_agree_with_oracle(table, host_pool_copy, lens):
    """Group kernel, paged form (Llama-3-8B shapes): with the host's copy of the pool addresses K/V arrive 128-byte swizzled
    through tiled boxes (runs of 16 consecutive slots) or tile::gather4 (anything else, ragged last tiles included) and the
    attention runs on the tensor cores; without it (or with a stale copy) the linear row gather + CUDA-core loop runs.
    All against the oracle, batch 1 and a ragged batch of 3."""
    d, indptr, indices, positions, cos_sin, kp, vp, want_o, want_r = _paged_case(lens, seed=808, table=table, shape=S8)
    out, rout, kpool, vpool = _run_paged_cabi(d, indptr, indices, positions, cos_sin, host_pool_copy=host_pool_copy, shape=S8)
    assert torch.equal(rout.cpu(), want_r)
    assert close(out, want_o)
    assert close_k(kpool, kp) and close(vpool, vp)


GQA_BATCH_CASES = {
    "8b_bs8_kv1k": (S8, [1024] * 8, "random"),
    "8b_bs8_ragged": (S8, [700, 0, 33, 2100, 1, 16, 640, 5], "runs"),
    "8b_bs9_two_chunks": (S8, [100, 257, 0, 31, 1000, 48, 16, 15, 300], "random"),
    "8b_bs16": (S8, [64 + 37 * i for i in range(16)], "sequential"),
    "8b_bs3_all_empty": (S8, [0, 0, 0], "random"),
    "8b_bs3_long": (S8, [16384, 5, 3000], "random"),
    "70b_bs5": (S70, [300, 64, 0, 1000, 17], "random"),
    "70b_shard4_bs3": (O.LayerShape(8192, 16, 2), [500, 0, 129], "runs"),
    "70b_shard8_bs4": (O.LayerShape(8192, 8, 1), [77, 1024, 0, 3], "random"),
}


@pytest.mark.parametrize("case", sorted(GQA_BATCH_CASES))
def test_gqa_batched_weights_once_kernel_vs_oracle(case):
    """Grouped-query shapes at batch >= 2 run llama_decoder_gqa_batch_kernel.cuh: weights streamed once per chunk of 8 requests,
    requests on the N dimension of the MMA, KV rows of the chunk cut into equal segments over the group's CTAs.  Against the
    oracle: full and ragged chunks, two chunks, empty requests (new token only), a long request next to short ones, Llama-2-70B
    shapes (hidden 8192: G = 8, 1024 columns per CTA) and its 4- / 8-way shards (G = 32 / 64: one RoPE pair per CTA and slot)."""
    shape, lens, table = GQA_BATCH_CASES[case]
    d, indptr, indices, positions, cos_sin, kp, vp, want_o, want_r = _paged_case(lens, seed=len(lens) * 7 + shape.hidden, table=table, shape=shape)
    out, rout, kpool, vpool = _run_paged_cabi(d, indptr, indices, positions, cos_sin, shape=shape)
    assert torch.equal(rout.cpu(), want_r)
    assert close(out, want_o)
    assert close_k(kpool, kp) and close(vpool, vp)


def test_gqa_batched_weights_once_kernel_matches_per_request_launch_inplace_and_repeatable():
    """Same call three ways: weights-once kernel, weights-once kernel with residual_out == residual_in (the chunk's CTA 0 rewrites
    the residual once every CTA is past phase 0), and the group kernel launched per request (CF_FLAG_PER_REQUEST).  The
    weights-once kernel has no atomics: 20 launches on the same workspace give bit-identical outputs and a clean status word."""
    import cabi_torch as CT
    from clusterfusion_b200 import cabi
    lens = [333, 0, 2100, 64, 1, 900]
    d, indptr, indices, positions, cos_sin, kp, vp, want_o, want_r = _paged_case(lens, seed=606, table="random", shape=S8)
    out, rout, kpool, vpool = _run_paged_cabi(d, indptr, indices, positions, cos_sin, shape=S8)
    out_i, rout_i, kpool_i, vpool_i = _run_paged_cabi(d, indptr, indices, positions, cos_sin, shape=S8, inplace=True)
    out_p, rout_p, kpool_p, vpool_p = _run_paged_cabi(d, indptr, indices, positions, cos_sin, shape=S8, flags=cabi.CF_FLAG_PER_REQUEST)
    assert torch.equal(rout.cpu(), want_r) and torch.equal(rout_i.cpu(), want_r) and torch.equal(rout_p.cpu(), want_r)
    assert torch.equal(out, out_i) and torch.equal(kpool, kpool_i) and torch.equal(vpool, vpool_i)
    assert close(out, want_o) and close(out_p, want_o)
    assert close(vpool, vpool_p) and close_k(kpool, kpool_p)     # (different fp32 summation orders: K-split vs row-split)
    # fp32 output (CF_FLAG_OUT_FP32_PARTIAL): the same sums before the final rounding
    out_f, _, _, _ = _run_paged_cabi(d, indptr, indices, positions, cos_sin, shape=S8, flags=cabi.CF_FLAG_OUT_FP32_PARTIAL, fp32_out=True)
    assert out_f.dtype == torch.float32 and torch.equal(out_f.half(), out)
    for _ in range(20):
        o2, _, k2, _ = _run_paged_cabi(d, indptr, indices, positions, cos_sin, shape=S8)
        assert torch.equal(o2, out) and torch.equal(k2, kpool)
    assert cabi.workspace_status(CT.workspace(4096, len(lens), out.device).data_ptr()) == 0


@pytest.mark.parametrize("shape_name", ["llama2_7b", "llama3_8b"])
def test_batched_kernels_random_ragged_batches_vs_oracle(shape_name):
    """The batched kernels cut the chunk's concatenated KV rows into equal segments over the CTAs of a head / group.  Random ragged
    batches (2..9 requests, lengths 0..400 with many empty and one-row requests, so that segment boundaries fall everywhere:
    inside tiles, on request boundaries, ranks with no rows at all) against the oracle, MHA (chunks of 4 / 8) and grouped-query."""
    shape = S7 if shape_name == "llama2_7b" else S8
    g = torch.Generator().manual_seed(20260)
    for case in range(10):
        bs = int(torch.randint(2, 10, (1,), generator=g))
        kinds = torch.randint(0, 4, (bs,), generator=g)
        lens = [0 if k == 0 else 1 if k == 1 else int(torch.randint(2, 400, (1,), generator=g)) for k in kinds.tolist()]
        d, indptr, indices, positions, cos_sin, kp, vp, want_o, want_r = _paged_case(lens, seed=7000 + case, table="random", shape=shape)
        out, rout, kpool, vpool = _run_paged_cabi(d, indptr, indices, positions, cos_sin, shape=shape)
        assert torch.equal(rout.cpu(), want_r), (case, lens)
        assert close(out, want_o), (case, lens)
        assert close_k(kpool, kp) and close(vpool, vp), (case, lens)


def test_gqa_batched_kernel_on_a_workspace_sized_for_a_larger_batch():
    """The public operator keeps one workspace per (device, stream, hidden) sized for the largest batch seen: a batch of 9 first
    (two chunks), then a batch of 3 on the same, larger workspace (CfLlamaArgs.workspace_batch > batch) -- both against the oracle."""
    import clusterfusion
    for lens, seed in (([40, 0, 300, 17, 64, 1, 128, 33, 250], 91), ([200, 31, 0], 92)):
        d, indptr, indices, positions, cos_sin, kp, vp, want_o, want_r = _paged_case(lens, seed=seed, table="random", shape=S8)
        c = cuda(d)
        kpool, vpool = c["k_cache"].clone(), c["v_cache"].clone()
        kptrs = torch.tensor([kpool.data_ptr()], dtype=torch.uint64).cuda()
        vptrs = torch.tensor([vpool.data_ptr()], dtype=torch.uint64).cuda()
        out = torch.full((len(lens), 4096), float("nan"), dtype=torch.float16, device="cuda"); rout = torch.empty_like(out)
        clusterfusion.llama_decoder_layer_batch_decode_sglang(out, rout, c["x"], c["residual"], c["weight_qkv"], c["weight_o"],
                                                              indptr.cuda(), indices.cuda(), kptrs, vptrs, 0, c["rms_w"], 1e-5,
                                                              positions.cuda(), cos_sin.cuda())
        torch.cuda.synchronize()
        assert torch.equal(rout.cpu(), want_r)
        assert close(out, want_o)
        assert close_k(kpool, kp) and close(vpool, vp)
    assert clusterfusion.workspace_status() == 0


@pytest.mark.parametrize("table", ["random", "sequential"])
def test_paged_form_kv16384_bs1_vs_oracle(table):
    """north_star's form at north_star's size: 15-argument paged call, batch 1, kv_len 16384 (Llama-2-7B), through the public
    operator (which passes the host copy of the pool addresses, so the tiled / gather4 paths run), oracle run in full."""
    import clusterfusion
    d, indptr, indices, positions, cos_sin, kp, vp, want_o, want_r = _paged_case([16384], seed=1616, table=table)
    c = cuda(d)
    kpool, vpool = c["k_cache"].clone(), c["v_cache"].clone()
    kptrs = torch.tensor([kpool.data_ptr()], dtype=torch.uint64).cuda()
    vptrs = torch.tensor([vpool.data_ptr()], dtype=torch.uint64).cuda()
    out = torch.full((1, 4096), float("nan"), dtype=torch.float16, device="cuda"); rout = torch.empty_like(out)
    clusterfusion.llama_decoder_layer(out, rout, c["x"], c["residual"], c["weight_qkv"], c["weight_o"], indptr.cuda(),
                                      indices.cuda(), kptrs, vptrs, 0, c["rms_w"], 1e-5, positions.cuda(), cos_sin.cuda())
    torch.cuda.synchronize()
    assert torch.equal(rout.cpu(), want_r)
    assert close(out, want_o)
    assert close_k(kpool, kp) and close(vpool, vp)


@pytest.mark.parametrize("per_request", [False, True])
def test_paged_form_ragged_bs4_long_context_vs_oracle(per_request):
    """Ragged batch of 4 with two requests at kv 16384: the batched kernel (weights once per chunk) and the one-cluster-per-
    (request, head) launch of the reference's grid shape, both against the oracle."""
    from clusterfusion_b200 import cabi
    lens = [16384, 5000, 16384, 1]
    d, indptr, indices, positions, cos_sin, kp, vp, want_o, want_r = _paged_case(lens, seed=44, table="random")
    out, rout, kpool, vpool = _run_paged_cabi(d, indptr, indices, positions, cos_sin,
                                              flags=cabi.CF_FLAG_PER_REQUEST if per_request else 0)
    assert torch.equal(rout.cpu(), want_r)
    assert close(out, want_o)
    assert close_k(kpool, kp) and close(vpool, vp)


def test_chat_operator_kv65536_vs_oracle():
    """8-argument form at the top of BASELINE config 3's sweep (kv 65536: 1 GB of K/V, 1024 tiles per CTA), oracle run in full."""
    import clusterfusion
    kv = 65536
    d = O.make_inputs(S7, kv, seed=65, layout="chat")
    want_o, want_k, want_v = O.chat_layer(d["x"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"],
                                          d["rms_w"], d["cos"], d["sin"], n_heads=32, eps=1e-6, mode="eager")
    c = cuda(d)
    o, k, v = clusterfusion.llama_decoder_layer(c["x"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"],
                                                c["rms_w"], c["cos"], c["sin"])
    torch.cuda.synchronize()
    assert close(o, want_o) and close(v, want_v) and close_k(k, want_k, pairing="gptj")


def test_chat_form_growing_cache_64_tokens_like_the_reference_chat_loop():
    """The 8-argument form driven the way /root/reference/chat/llama/model.py:355-374 drives it: `k_cache` / `v_cache` are
    views `[:start_pos]` of larger per-layer buffers, kv_len grows by one every token, the caller appends the returned k / v.
    64 tokens x 2 layers against the oracle fed with the same growing cache; and the host path must be in steady state after
    the first token: ZERO cuTensorMapEncodeTiled calls for tokens 2..64 (maps are keyed on the base pointer, not kv_len)."""
    import clusterfusion
    n_layers, start, n_tok, max_seq = 2, 200, 64, 512
    g = torch.Generator().manual_seed(2024)
    layers = []
    for li in range(n_layers):
        d = O.make_inputs(S7, start, seed=900 + li, layout="chat")
        ck = torch.zeros(max_seq, 4096, dtype=torch.float16); cv = torch.zeros_like(ck)
        ck[:start] = d["k_cache"]; cv[:start] = d["v_cache"]
        layers.append(dict(w_qkv=d["weight_qkv"], w_o=d["weight_o"], rms=d["rms_w"], ck=ck, cv=cv))
    dev = [{k: v.cuda() for k, v in l.items()} for l in layers]
    x0 = torch.randn(1, 1, 4096, generator=g).half()
    h_ref, h_dev = x0.clone(), x0.cuda()
    worst = 0.0
    enc_after_first = None
    for t in range(n_tok):
        pos = start + t
        ang = O.rope_angles(pos)
        cos = torch.repeat_interleave(ang.cos(), 2).view(1, 128).contiguous()
        sin = torch.repeat_interleave(ang.sin(), 2).view(1, 128).contiguous()
        cos_d, sin_d = cos.cuda(), sin.cuda()
        for l, ld in zip(layers, dev):
            # oracle, fed with what the DEVICE path consumed (so a 1-ulp difference in token t cannot drift into token t+1)
            want_o, want_k, want_v = O.chat_layer(h_dev.cpu().view(1, 4096), l["w_qkv"], l["w_o"], ld["ck"][:pos].cpu(), ld["cv"][:pos].cpu(),
                                                  l["rms"], cos, sin, n_heads=32, eps=1e-6, mode="eager")
            o, xk, xv = clusterfusion.llama_decoder_layer(h_dev, ld["w_qkv"], ld["w_o"], ld["ck"][:pos], ld["cv"][:pos],
                                                          ld["rms"], cos_d, sin_d)
            ld["ck"][pos:pos + 1] = xk.view(1, 4096)          # model.py:371-372
            ld["cv"][pos:pos + 1] = xv.view(1, 4096)
            assert close(o, want_o) and close(xv, want_v) and close_k(xk, want_k, name=f"k[t={t}]", pairing="gptj")
            worst = max(worst, float((o.float().cpu() - want_o.float()).abs().max()))
            h_dev = (h_dev + o.view(1, 1, 4096)) * 0.5          # keep activations O(1) over 128 layer calls
        if t == 0:
            torch.cuda.synchronize()
            enc_after_first = clusterfusion.tensor_map_encodes()
    torch.cuda.synchronize()
    print(f"growing cache: {n_tok} tokens x {n_layers} layers, worst |o - oracle| = {worst:.2e}, "
          f"tensor-map encodes after token 1: {clusterfusion.tensor_map_encodes() - enc_after_first}")
    assert clusterfusion.tensor_map_encodes() == enc_after_first


def test_workspace_status_word_round_trip():
    """cf_workspace_status / cf_workspace_clear_status (C ABI): the sticky error word of a workspace reads 0 after normal
    launches, reads back a planted value, and clears."""
    import cabi_torch as CT
    from clusterfusion_b200 import cabi
    d = O.make_inputs(S7, 64, seed=5, layout="chat")
    c = cuda(d)
    CT.chat(c["x"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], c["cos"], c["sin"])
    ws = CT.workspace(4096, 1, c["x"].device)
    assert cabi.workspace_status(ws.data_ptr(), CT.stream_handle()) == 0
    ws.view(torch.int32)[2] = 1
    assert cabi.workspace_status(ws.data_ptr(), CT.stream_handle()) == 1
    cabi.workspace_clear_status(ws.data_ptr(), CT.stream_handle())
    assert cabi.workspace_status(ws.data_ptr(), CT.stream_handle()) == 0


def test_flag_in_data_epoch_survives_the_32_bit_wrap():
    """The launch epoch in the workspace header is a 32-bit counter bumped once per launch (10-100k launches/s in a serving
    process: it wraps within hours).  Preset it to 0xFFFFFFFE and run launches across the wrap through both users of the
    flag-in-data words -- the group kernel and the MHA kernel's CF_FLAG_LL_OUT path: every launch must still match the oracle
    (consecutive flags distinct, never 0) and no poll may time out."""
    import cabi_torch as ct
    from clusterfusion_b200 import cabi
    for shape, flags in ((S8, 0), (S7, cabi.CF_FLAG_LL_OUT)):
        d = O.make_inputs(shape, 300, seed=9, layout="sglang")
        want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], 1e-5,
                              d["cos"], d["sin"], n_heads=shape.n_heads, n_kv_heads=shape.n_kv_heads, mode="eager")
        c = cuda(d)
        ws = ct.workspace(shape.hidden, 1, c["x"].device)
        torch.cuda.synchronize()
        ws.view(torch.int32)[0] = -2                      # 0xFFFFFFFE
        for it in range(5):
            o, r, k, v = ct.sglang(c["x"], c["residual"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], 1e-5,
                                   c["cos"], c["sin"], n_heads=shape.n_heads, n_kv_heads=shape.n_kv_heads, flags=flags)
            torch.cuda.synchronize()
            assert close(o, want[0]), (shape, it)
        hdr = ws[:16].view(torch.int32).cpu()
        assert int(hdr[0]) == 3 and int(hdr[2]) == 0, hdr          # 0xFFFFFFFE + 5 launches = 3 (mod 2^32); no poll time-out


def test_soak_10000_launches_as_the_reference_test_runs_them():
    """/root/reference/tests/test_llama.py:17-22, :143-157 launches the 10-argument operator `test_run = 10000` times on the same
    inputs (kv_len 4096, a fresh residual clone per launch) and asserts nothing.  Same loop here, with the asserts: every 500th
    output is kept and all of them must match the oracle, agree with the first to 1 fp16 ulp (the cross-head fp32 sum is the
    only order-dependent step), the in-place residual must be right on the last launch, and no workspace word may be left set."""
    import clusterfusion
    d = O.make_inputs(S7, 4096, seed=4242, layout="sglang")
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], 1e-6,
                          d["cos"], d["sin"], n_heads=32, mode="eager")
    c = cuda(d)
    kept = []
    for i in range(10000):
        tmp_residual = c["residual"].clone()
        o, r, k, v = clusterfusion.llama_decoder_layer_sglang(c["x"], tmp_residual, c["weight_qkv"], c["weight_o"], c["k_cache"],
                                                              c["v_cache"], c["rms_w"], 1e-6, c["cos"], c["sin"])
        if i % 500 == 0 or i == 9999:
            kept.append(o)
    torch.cuda.synchronize()
    assert torch.equal(r.cpu(), want[1])
    ref = kept[0].float()
    ulp = torch.maximum(ref.abs() * 2 ** -10, torch.full_like(ref, 2 ** -24))
    for o in kept:
        assert close(o, want[0])
        assert bool(((o.float() - ref).abs() <= ulp).all())
    assert close(v, want[3]) and close_k(k, want[2])
    assert clusterfusion.workspace_status() == 0


def test_first_call_inside_stream_capture_and_debug_status_mode():
    """The public operators keep one workspace per (device, stream, hidden).  If the FIRST call on a stream happens inside a
    CUDA-graph capture, the workspace must be allocated and zeroed OUTSIDE the graph (relaxed capture mode, private non-blocking
    stream): a memset recorded into the graph would re-zero the workspace on every replay and the buffer would die with the
    graph's pool.  Replays must match the oracle, the epoch must advance once per replayed launch, and the debug mode
    set_check_status(True) (read the error word back after every launch) must stay silent."""
    import clusterfusion
    d = O.make_inputs(S8, 500, seed=77, layout="sglang", theta=500000.0)
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], 1e-5,
                          d["cos"], d["sin"], n_heads=32, n_kv_heads=8, mode="eager")
    c = cuda(d)
    res = c["residual"].clone()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):          # first use of stream `s` by the module: workspace created during capture
        o, r, k, v = clusterfusion.llama_decoder_layer_sglang(c["x"], res, c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"],
                                                              c["rms_w"], 1e-5, c["cos"], c["sin"])
    for _ in range(3):
        res.copy_(c["residual"])
        g.replay()
        torch.cuda.synchronize()
        assert close(o, want[0]) and close(v, want[3]) and close_k(k, want[2]) and torch.equal(res.cpu(), want[1])
    assert clusterfusion.workspace_status() == 0
    clusterfusion.set_check_status(True)
    try:
        o2, _, _, _ = clusterfusion.llama_decoder_layer_sglang(c["x"], c["residual"].clone(), c["weight_qkv"], c["weight_o"], c["k_cache"],
                                                               c["v_cache"], c["rms_w"], 1e-5, c["cos"], c["sin"])
        assert close(o2, want[0])
    finally:
        clusterfusion.set_check_status(False)


def test_aliasing_outputs_are_rejected():
    """output / residual_output overlapping input / residual (other than residual_output IS residual) would be silently corrupted
    by the in-place epilogues: the operator must raise instead."""
    import clusterfusion
    d = O.make_inputs(S7, 10, seed=3, layout="sglang", bs=2)
    c = cuda(d)
    pool = torch.zeros(32, 4096, dtype=torch.float16, device="cuda")
    kp = torch.tensor([pool.data_ptr()], dtype=torch.uint64).cuda()
    indptr = torch.tensor([0, 3, 6], dtype=torch.int32).cuda(); indices = torch.arange(6, dtype=torch.int32).cuda()
    pos = torch.tensor([2, 2], dtype=torch.int64).cuda(); cs = torch.rand(8, 128).cuda()
    out = torch.empty_like(c["x"]); rout = torch.empty_like(c["x"])
    args = lambda o_, ro_, x_, r_: (o_, ro_, x_, r_, c["weight_qkv"], c["weight_o"], indptr, indices, kp, kp, 0, c["rms_w"], 1e-5, pos, cs)
    with pytest.raises(RuntimeError):
        clusterfusion.llama_decoder_layer_batch_decode_sglang(*args(c["x"], rout, c["x"], c["residual"]))          # output is input
    with pytest.raises(RuntimeError):
        clusterfusion.llama_decoder_layer_batch_decode_sglang(*args(out, c["x"], c["x"], c["residual"]))           # residual_output is input
    both = torch.empty(3, 4096, dtype=torch.float16, device="cuda")
    with pytest.raises(RuntimeError):
        clusterfusion.llama_decoder_layer_batch_decode_sglang(*args(both[:2], both[1:], c["x"], c["residual"]))    # outputs overlap each other
    clusterfusion.llama_decoder_layer_batch_decode_sglang(*args(out, c["residual"].clone(), c["x"], c["residual"]))
    torch.cuda.synchronize()

"""GPU parity against the REFERENCE'S OWN KERNELS (B200 box, -m gpu).

oracle/_ref holds the unmodified reference extension (xinhao-luo/ClusterFusion include/pybind.cpp + include/H100/*)
compiled for sm_100a by oracle/build_ref.sh.  It is the strongest pin available for the oracle and for our kernels:
the same seeded inputs go through (1) the reference kernel, (2) our kernel, (3) the CPU oracle.

What is asserted:
  * our output is within rtol = atol = 1e-3 of the oracle (the north-star bar) -- as in test_gpu_parity.py,
  * the reference kernel's output is within its own documented tolerance of the oracle (5e-2 on the output, 1e-2 on
    k / v: /root/reference/tests/test_llama_tilelang.py:100; the 10-argument kernel updates `residual` in place while
    other clusters still read it, SURVEY.md Q6 -- on B200 that race fires on some launches, so its best of 6 launches
    is the one compared), i.e. the oracle really states what the reference computes,
  * ours is at least as close to the fp32 oracle as the reference kernel is (the reference rounds partial sums to fp16
    and sums heads with fp16 atomics).

Skipped when oracle/_ref was not built (it needs /root/reference at build time).
"""
import importlib.util
import sys
from pathlib import Path

import pytest
import torch

from oracle import llama_oracle as O

from parity_helpers import close_k  # noqa: E402

pytestmark = pytest.mark.gpu

REF_DIR = Path(__file__).resolve().parent.parent / "oracle" / "_ref"
S7 = O.LayerShape(4096, 32, 32)


def load_ref():
    sos = sorted(REF_DIR.glob("_clusterfusion_ref*.so"))
    if not sos:
        pytest.skip("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    spec = importlib.util.spec_from_file_location("_clusterfusion_ref", sos[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="module")
def ref():
    return load_ref()


def err(a, b):
    return float((a.float().cpu().reshape(-1) - b.float().cpu().reshape(-1)).abs().max())


@pytest.mark.parametrize("kv_len", [37, 1024, 4096])
def test_chat_form_vs_reference_kernel(ref, kv_len):
    import clusterfusion
    d = O.make_inputs(S7, kv_len, seed=kv_len, layout="chat")
    want = O.chat_layer(d["x"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], d["cos"], d["sin"],
                        n_heads=32, eps=1e-6, mode="eager")
    truth = O.chat_layer(d["x"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], d["cos"], d["sin"],
                         n_heads=32, eps=1e-6, mode="fp32")
    c = {k: v.cuda() for k, v in d.items()}
    args = (c["x"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], c["cos"], c["sin"])
    r_o, r_k, r_v = ref.llama_decoder_layer(*args)
    torch.cuda.synchronize()
    o, k, v = clusterfusion.llama_decoder_layer(*args)
    torch.cuda.synchronize()
    # ours vs oracle: the north-star tolerance
    assert torch.allclose(o.float().cpu(), want[0].float(), rtol=1e-3, atol=1e-3)
    assert torch.allclose(v.float().cpu(), want[2].float(), rtol=1e-3, atol=1e-3)
    assert close_k(k, want[1], pairing="gptj")            # K after RoPE: stated in ulps of the pair magnitude (parity_helpers.close_k)
    # reference kernel vs oracle: the reference's own tolerance
    assert err(r_o, want[0]) < 5e-2 and err(r_k, want[1]) < 1e-2 and err(r_v, want[2]) < 1e-2
    # ours vs the reference kernel directly
    assert err(o, r_o) < 5e-2 and err(k, r_k) < 1e-2 and err(v, r_v) < 1e-2
    # and we are at least as accurate as the reference against the unrounded fp32 statement
    assert err(o, truth[0]) <= err(r_o, truth[0]) + 1e-4
    print(f"kv={kv_len}: |ours-oracle|={err(o, want[0]):.2e} |ref-oracle|={err(r_o, want[0]):.2e} |ours-ref|={err(o, r_o):.2e}")


@pytest.mark.parametrize("kv_len", [256, 4096])
def test_sglang_form_vs_reference_kernel(ref, kv_len):
    import clusterfusion
    d = O.make_inputs(S7, kv_len, seed=100 + kv_len, layout="sglang")
    eps = 1e-6
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], eps,
                          d["cos"], d["sin"], n_heads=32, mode="eager")
    c = {k: v.cuda() for k, v in d.items()}
    cos128 = torch.cat([c["cos"], c["cos"]]).contiguous()       # tests/test_llama.py:138-139 passes cat([cos, cos])
    sin128 = torch.cat([c["sin"], c["sin"]]).contiguous()
    # The reference kernel adds `input` into `residual` IN PLACE while other clusters may still be reading it
    # (kernel_sglang.cuh:100-105).  On B200 the race fires on some launches (observed: first launches after
    # a module load give |err| 0.07-0.2, steady-state launches 1e-3), so the reference gets 6 attempts and its best one
    # is compared; the number of racy launches is printed.
    runs = []
    for _ in range(6):
        r_res = c["residual"].clone()
        r = ref.llama_decoder_layer_sglang(c["x"], r_res, c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"],
                                           c["rms_w"], eps, cos128, sin128)
        torch.cuda.synchronize()
        runs.append(r)
    errs = [err(r[0], want[0]) for r in runs]
    print(f"reference sglang kernel, 6 launches: max-abs errors {[round(e, 4) for e in errs]}")
    r_o, r_res_out, r_k, r_v = runs[errs.index(min(errs))]
    res = c["residual"].clone()
    o, res_out, k, v = clusterfusion.llama_decoder_layer_sglang(c["x"], res, c["weight_qkv"], c["weight_o"], c["k_cache"],
                                                                c["v_cache"], c["rms_w"], eps, cos128, sin128)
    torch.cuda.synchronize()
    assert torch.allclose(o.float().cpu(), want[0].float(), rtol=1e-3, atol=1e-3)
    assert torch.equal(res_out.cpu(), want[1])
    print(f"kv={kv_len}: |ours-oracle|={err(o, want[0]):.2e} |ref-oracle|={err(r_o, want[0]):.2e} |ours-ref|={err(o, r_o):.2e} "
          f"ref k/v err {err(r_k, want[2]):.2e} {err(r_v, want[3]):.2e}")
    assert err(r_o, want[0]) < 5e-2 and err(o, r_o) < 5e-2
    assert err(r_k, want[2]) < 1e-2 and err(r_v, want[3]) < 1e-2 and err(k, r_k) < 1e-2 and err(v, r_v) < 1e-2
    print(f"kv={kv_len}: |ours-oracle|={err(o, want[0]):.2e} |ref-oracle|={err(r_o, want[0]):.2e} |ours-ref|={err(o, r_o):.2e}")


def test_paged_form_vs_reference_kernel(ref):
    import clusterfusion
    bs, L, layer_id, n_slots, eps = 2, 3, 1, 700, 1e-5
    lens = [300, 129]
    d = O.make_inputs(S7, 1, seed=7, layout="sglang", bs=bs)
    g = torch.Generator().manual_seed(3)
    pools_k = [torch.randn(n_slots, 4096, generator=g).half() for _ in range(L)]
    pools_v = [torch.randn(n_slots, 4096, generator=g).half() for _ in range(L)]
    perm = torch.randperm(n_slots, generator=g).int()
    indptr = torch.tensor([0, lens[0] + 1, lens[0] + lens[1] + 2], dtype=torch.int32)
    indices = perm[: int(indptr[-1])].contiguous()
    positions = torch.tensor(lens, dtype=torch.int64)
    inv = 1.0 / (10000.0 ** (torch.arange(0, 128, 2).float() / 128))
    ang = torch.outer(torch.arange(512).float(), inv)
    cos_sin = torch.cat([ang.cos(), ang.sin()], 1).contiguous()
    kp, vp = pools_k[layer_id].clone(), pools_v[layer_id].clone()
    want_o, want_r = O.paged_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], indptr, indices, kp, vp, d["rms_w"],
                                   eps, positions, cos_sin, n_heads=32, mode="eager")

    def run(fn):
        pk = [t.cuda() for t in pools_k]
        pv = [t.cuda() for t in pools_v]
        kptrs = torch.tensor([t.data_ptr() for t in pk], dtype=torch.uint64).cuda()
        vptrs = torch.tensor([t.data_ptr() for t in pv], dtype=torch.uint64).cuda()
        out = torch.zeros(bs, 4096, dtype=torch.float16, device="cuda")
        rout = torch.zeros(bs, 4096, dtype=torch.float16, device="cuda")
        fn(out, rout, d["x"].cuda(), d["residual"].cuda(), d["weight_qkv"].cuda(), d["weight_o"].cuda(), indptr.cuda(),
           indices.cuda(), kptrs, vptrs, layer_id, d["rms_w"].cuda(), eps, positions.cuda(), cos_sin.cuda())
        torch.cuda.synchronize()
        return out.cpu(), rout.cpu(), pk[layer_id].cpu(), pv[layer_id].cpu()

    o, r, k_pool, v_pool = run(clusterfusion.llama_decoder_layer_batch_decode_sglang)
    assert torch.allclose(o.float(), want_o.float(), rtol=1e-3, atol=1e-3)
    assert torch.equal(r, want_r)
    assert close_k(k_pool, kp)
    assert torch.allclose(v_pool.float(), vp.float(), rtol=1e-3, atol=1e-3)
    # The reference kernel zeroes its slice of `output` inside the kernel with no grid-wide ordering against the other
    # clusters' atomicAdds (kernel_batch_sglang.cuh:608-610 vs :643): a cluster that zeroes late wipes what
    # earlier clusters added.  On B200 its output is wrong on every launch (observed max-abs error 0.19-0.34) and the K/V
    # rows it appends to the pool are off as well (up to 1.6), so this kernel cannot serve as a pin on B200: only its
    # race-free `residual_output` is asserted, the rest is printed.  (The 8- and 10-argument reference kernels above DO
    # agree with the oracle, and the paged oracle is the 10-argument oracle applied per request.)
    runs = [run(ref.llama_decoder_layer_batch_decode_sglang) for _ in range(3)]
    print("reference paged kernel, 3 launches: max-abs errors vs oracle  output",
          [round(err(x[0], want_o), 4) for x in runs], " k pool", [round(err(x[2], kp), 4) for x in runs],
          " v pool", [round(err(x[3], vp), 4) for x in runs], "(not asserted)")
    for r_o, r_r, r_k_pool, r_v_pool in runs:
        assert torch.equal(r_r, want_r)

def test_rmsnorm_vs_reference_kernel(ref):
    """The reference's standalone op is fixed at 64 x 8192 (include/H100/norm/config.h:1-2)."""
    import clusterfusion
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 8192, generator=g).half()
    w = torch.randn(8192, generator=g).half()
    want = O.rmsnorm_op(x, w, 1e-6)
    r = ref.rmsnorm(x.cuda(), w.cuda())
    o = clusterfusion.rmsnorm(x.cuda(), w.cuda())
    torch.cuda.synchronize()
    print(f"rmsnorm: |ours-oracle|={err(o, want):.2e} |ref-oracle|={err(r, want):.2e} |ours-ref|={err(o, r):.2e}")
    assert err(o, want) < 4e-3 and err(r, want) < 4e-3 and err(o, r) < 4e-3        # values up to ~16: 1 fp16 ulp = 7.8e-3 / 2


def test_reference_deepseek_kernel_vs_oracle(ref):
    """The reference's DeepSeek-MLA kernel (fixed at hidden 2048 / 16 heads / SEQ_LEN 4096, no test of its own upstream) next to
    ours and to the oracle on the same inputs.  On B200 its output is not reproducible from launch to launch and is far from
    the oracle (tools/deepseek_probe.py measured |diff| ~ 33 on outputs of magnitude 2), so the oracle stays UNPINNED for this
    op: only our kernel is asserted, the reference's numbers are printed for the record."""
    import clusterfusion
    from oracle import deepseek_oracle as D
    d = D.make_inputs(4096, seed=4096, out_gain=2.4)
    keys = ("x", "w_q_nope", "w_q_pe", "w_uk", "w_kv", "w_k_pe", "w_uv", "w_o", "ckv_cache", "rms_in_w", "rms_ckv_w", "cos", "sin")
    c = [d[k].cuda() for k in keys]
    want, _, _ = D.deepseek_layer(**d)
    ours = clusterfusion.deepseek_decoder_layer(*c)
    torch.cuda.synchronize()
    assert torch.allclose(ours.float().cpu(), want.float(), rtol=1e-3, atol=1e-3)
    ours_err = err(ours, want)
    try:                                                  # nothing of the reference kernel's behaviour on B200 is relied upon
        errs = []
        for _ in range(4):
            o = ref.deepseek_decoder_layer(*c)
            torch.cuda.synchronize()
            errs.append(err(o, want))
        print(f"reference DeepSeek kernel vs oracle: {errs}; ours vs oracle: {ours_err}")
    except Exception as e:                                # noqa: BLE001 -- recorded, not asserted
        print(f"reference DeepSeek kernel failed on this GPU: {e!r}; ours vs oracle: {ours_err}")

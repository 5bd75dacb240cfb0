"""GPU probe for the DeepSeek-MLA half-layer: parity against the CPU oracle over a range of cache lengths, repeated and
PDL-chained launches, the reference's own kernel (oracle/_ref, seq_len 4096 only) and a CUDA-graph timing.
    python tools/deepseek_probe.py            (run under `timeout`)"""
import importlib.util
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, ".")
import clusterfusion_b200 as cf
from oracle import deepseek_oracle as D

dev = torch.device("cuda", 0)
KEYS = ("x", "w_q_nope", "w_q_pe", "w_uk", "w_kv", "w_k_pe", "w_uv", "w_o", "ckv_cache", "rms_in_w", "rms_ckv_w", "cos", "sin")


def err(a, b):
    return float((a.float().cpu().reshape(-1) - b.float().cpu().reshape(-1)).abs().max())


res = []
for S, gain in ((1, 0.75), (2, 0.75), (17, 1.0), (33, 1.0), (300, 1.5), (4096, 2.4), (4097, 2.4), (20000, 3.0)):
    d = D.make_inputs(S, seed=S, out_gain=gain)
    c = [d[k].to(dev) for k in KEYS]
    for rope in (False, True):
        want, ckv, kpe = D.deepseek_layer(**d, rope_scores=rope)
        got, gckv, gkpe = cf.deepseek_decoder_layer_ex(*c, rope)
        torch.cuda.synchronize()
        got2, _, _ = cf.deepseek_decoder_layer_ex(*c, rope)
        torch.cuda.synchronize()
        ok = torch.allclose(got.float().cpu(), want.float(), rtol=1e-3, atol=1e-3)
        row = dict(seq_len=S, rope_scores=rope, out_err=err(got, want), out_absmax=float(want.float().abs().max()), allclose_1e3=bool(ok),
                   repeat_err=err(got, got2), ckv_err=err(gckv, ckv), k_pe_err=err(gkpe, kpe))
        print(json.dumps(row), flush=True)
        res.append(row)

# timing: 16 calls with different weights per call inside one CUDA graph
for S in (4096, 16384, 65536):
    nl = 16
    L = []
    for i in range(nl):
        d = D.make_inputs(S, seed=i, out_gain=2.4) if i < 2 else None
        if d is not None:
            L.append([d[k].to(dev) for k in KEYS])
        else:
            L.append([t.clone() for t in L[i % 2]])
    for pdl in (False, True):
        cf.set_pdl(pdl)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for i in range(nl):
                cf.deepseek_decoder_layer(*L[i])
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            x = L[0][0]
            for i in range(nl):
                a = list(L[i]); a[0] = x
                x = cf.deepseek_decoder_layer(*a)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / (20 * nl))
        byts = 2 * (2048 * 2048 + 2048 * 1024 + 128 * 8192 + 2048 * 512 + 2048 * 64 + 512 * 2048 + 2048 * 2048) + S * 1152
        print(json.dumps(dict(seq_len=S, pdl=pdl, us_per_layer=round(best, 2), algorithmic_MB=round(byts / 1e6, 2),
                              GBps=round(byts / best / 1e3, 1), finite=bool(torch.isfinite(x.float()).all()))), flush=True)
    cf.set_pdl(False)
    del L

# the reference's own kernel (fixed SEQ_LEN 4096), if it was built
sos = sorted(Path("oracle/_ref").glob("_clusterfusion_ref*.so"))
if sos:
    spec = importlib.util.spec_from_file_location("_clusterfusion_ref", sos[0])
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    d = D.make_inputs(4096, seed=4096, out_gain=2.4)
    c = [d[k].to(dev) for k in KEYS]
    want, _, _ = D.deepseek_layer(**d)
    exact, _, _ = D.deepseek_layer(**d, mode="exact")
    ours = cf.deepseek_decoder_layer(*c)
    outs = []
    for _ in range(4):
        o = ref.deepseek_decoder_layer(*c)
        torch.cuda.synchronize()
        outs.append(o.clone())
    print(json.dumps(dict(reference_kernel=True, ref_vs_oracle=[err(o, want) for o in outs], ours_vs_oracle=err(ours, want),
                          ours_vs_exact=err(ours, exact), ref_vs_exact=err(outs[-1], exact), ref_vs_ours=err(outs[-1], ours),
                          out_absmax=float(want.float().abs().max()))), flush=True)
    # how long the reference's kernel takes on this GPU: CUDA events around its operator (3 device syncs and 9 tensor-map
    # encodes per call, deepseek_kernel_dispatch.cu:40-241) and the kernel alone from CUPTI
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ref.deepseek_decoder_layer(*c)
    e1.record()
    torch.cuda.synchronize()
    us_call, us_kernel = e0.elapsed_time(e1) * 1e3 / 20, None
    try:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(10):
                ref.deepseek_decoder_layer(*c)
            torch.cuda.synchronize()
        ks = [e for e in prof.events() if "DeepSeekDecoderLayerKernel" in e.name]
        if ks:
            us_kernel = sum(e.device_time for e in ks) / len(ks)
    except Exception as ex:                      # noqa: BLE001
        print("profiler unavailable:", repr(ex))
    print(json.dumps(dict(reference_kernel_time=True, seq_len=4096, us_per_call=round(us_call, 2),
                          us_kernel=None if us_kernel is None else round(us_kernel, 2))), flush=True)


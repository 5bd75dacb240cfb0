"""GPU probe (not product code): parity + timing of the DeepSeek-MLA half-layer for the kernel library named by CF_LIB_PATH
(default: the in-tree one), through the C ABI only -- used to A/B kernel variants built into build/variants/ on one box.
    CF_LIB_PATH=build/variants/libcf_ds_sw.so python tools/ds_ab.py [flags-to-or-in]"""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from clusterfusion_b200 import cabi          # noqa: E402
from oracle import deepseek_oracle as D      # noqa: E402

extra_flags = int(sys.argv[1], 0) if len(sys.argv) > 1 else 0
dev = torch.device("cuda", 0)
lib = cabi.load()
ws = torch.zeros(lib.cf_deepseek_workspace_bytes(), dtype=torch.uint8, device=dev)


def launch(c, out, seq_len, rope, pdl, stream):
    a = cabi.CfDeepseekArgs(flags=(cabi.CF_FLAG_PDL if pdl else 0) | (cabi.CF_DS_FLAG_ROPE_SCORES if rope else 0) | extra_flags,
                            hidden=D.HIDDEN, n_heads=D.N_HEADS, seq_len=seq_len, eps=1e-6, x=c["x"].data_ptr(),
                            w_q_nope=c["w_q_nope"].data_ptr(), w_q_pe=c["w_q_pe"].data_ptr(), w_uk=c["w_uk"].data_ptr(),
                            w_kv_nope=c["w_kv"].data_ptr(), w_k_pe=c["w_k_pe"].data_ptr(), w_uv=c["w_uv"].data_ptr(),
                            w_o=c["w_o"].data_ptr(), ckv_cache=c["ckv_cache"].data_ptr(), rms_input_w=c["rms_in_w"].data_ptr(),
                            rms_ckv_w=c["rms_ckv_w"].data_ptr(), cos=c["cos"].data_ptr(), sin=c["sin"].data_ptr(),
                            out=out.data_ptr(), workspace=ws.data_ptr())
    cabi.launch_deepseek(a, stream)


res = {"lib": os.environ.get("CF_LIB_PATH", "in-tree"), "flags": extra_flags, "parity": [], "timing": []}
for S, gain in ((1, 0.75), (2, 0.75), (34, 1.0), (300, 1.5), (4096, 2.4), (4097, 2.4), (9000, 3.0)):
    d = D.make_inputs(S, seed=S, out_gain=gain)
    c = {k: v.to(dev) for k, v in d.items()}
    for rope in (False, True):
        want, _, _ = D.deepseek_layer(**d, rope_scores=rope)
        outs = []
        for pdl in (False, True, True, False):
            o = torch.empty(1, D.HIDDEN, dtype=torch.float16, device=dev)
            launch(c, o, S, rope, pdl, torch.cuda.current_stream().cuda_stream)
            outs.append(o)
        torch.cuda.synchronize()
        errs = [float((o.float().cpu() - want.float()).abs().max()) for o in outs]
        ok = all(torch.allclose(o.float().cpu(), want.float(), rtol=1e-3, atol=1e-3) for o in outs)
        res["parity"].append({"seq_len": S, "rope": rope, "ok": bool(ok), "max_err": max(errs)})
        print(json.dumps(res["parity"][-1]), flush=True)

g = torch.Generator(device=dev).manual_seed(31)
r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).half()
for S in (4096, 16384):
    nl = 16
    L = [dict(x=None, w_q_nope=r(2048, 2048, sc=0.022), w_q_pe=r(2048, 1024, sc=0.022), w_uk=r(128, 8192, sc=0.088), w_kv=r(2048, 512, sc=0.022),
              w_k_pe=r(2048, 64, sc=0.022), w_uv=r(512, 2048, sc=0.1), w_o=r(2048, 2048, sc=0.05), ckv_cache=r(S, 576),
              rms_in_w=(1 + 0.1 * r(2048).float()).half(), rms_ckv_w=(1 + 0.1 * r(512).float()).half(),
              cos=torch.rand(64, generator=g, device=dev), sin=torch.rand(64, generator=g, device=dev),
              o=torch.empty(1, 2048, dtype=torch.float16, device=dev)) for _ in range(nl)]
    x = r(1, 2048)

    def chain(stream):
        h = x
        for lay in L:
            lay["x"] = h
            launch(lay, lay["o"], S, False, True, stream)
            h = lay["o"]
    s_ = torch.cuda.Stream()
    with torch.cuda.stream(s_):
        chain(s_.cuda_stream)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        chain(torch.cuda.current_stream().cuda_stream)
    for _ in range(5):
        gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        for _ in range(40):
            gr.replay()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / (40 * nl))
    res["timing"].append({"seq_len": S, "us_per_layer": round(best, 2)})
    print(json.dumps(res["timing"][-1]), flush=True)
    del gr, L
    torch.cuda.empty_cache()
print("SUMMARY", json.dumps(res))

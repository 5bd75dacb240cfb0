"""Debug helper (GPU): run each variant several times and report per-head mismatches vs the oracle."""
import sys, torch
sys.path.insert(0, ".")
from oracle import llama_oracle as O
import clusterfusion
sys.path.insert(0, "tests")
import cabi_torch as ct

S7 = O.LayerShape(4096, 32, 32)

def heads_bad(got, want, atol=1e-3, rtol=1e-3, per=128):
    g, w = got.float().cpu().reshape(-1), want.float().reshape(-1)
    bad = (g - w).abs() > atol + rtol * w.abs()
    hb = bad.view(-1, per).any(1).nonzero().flatten().tolist()
    return hb, float((g - w).abs().max())

for variant in ("chat", "sglang"):
    for kv in (0, 1, 31, 37, 256, 1024, 4096):
        d = O.make_inputs(S7, kv, seed=42 + kv, layout=variant)
        c = {k: v.cuda() for k, v in d.items()}
        if variant == "chat":
            want = O.chat_layer(d["x"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], d["cos"], d["sin"], n_heads=32, mode="eager")
        else:
            o, r, k, v = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], 1e-5, d["cos"], d["sin"], n_heads=32, mode="eager")
            want = (o, k, v)
        res = []
        for it in range(6):
            if variant == "chat":
                o, k, v = clusterfusion.llama_decoder_layer(c["x"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], c["cos"], c["sin"])
            else:
                o, r, k, v = ct.sglang(c["x"], c["residual"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], 1e-5, c["cos"], c["sin"], n_heads=32)
            torch.cuda.synchronize()
            ob, od = heads_bad(o, want[0], per=1024)
            kb, kd = heads_bad(k, want[1], atol=4e-3)
            vb, vd = heads_bad(v, want[2])
            res.append((ob, kb, vb, round(od, 4), round(kd, 4), round(vd, 4)))
        print(variant, kv, res, flush=True)

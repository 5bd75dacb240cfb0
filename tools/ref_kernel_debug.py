"""GPU probe: error statistics of the reference's own kernels (oracle/_ref) against the oracle, 10 runs each, to tell
arithmetic noise from races."""
import sys, importlib.util, glob, torch
sys.path.insert(0, ".")
from oracle import llama_oracle as O
so = glob.glob("oracle/_ref/_clusterfusion_ref*.so")[0]
spec = importlib.util.spec_from_file_location("_clusterfusion_ref", so); ref = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref)
import clusterfusion
S7 = O.LayerShape(4096, 32, 32)
def stats(tag, got, want):
    d = (got.float().cpu().reshape(-1) - want.float().reshape(-1)).abs()
    print(f"  {tag}: max {d.max():.3e} mean {d.mean():.3e} n>1e-2 {(d > 1e-2).sum().item()} n>5e-2 {(d > 5e-2).sum().item()} argmax {int(d.argmax())}")
for kv in (256, 4096):
    d = O.make_inputs(S7, kv, seed=100 + kv, layout="sglang")
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], 1e-6, d["cos"], d["sin"], n_heads=32, mode="eager")
    c = {k: v.cuda() for k, v in d.items()}
    cos128 = torch.cat([c["cos"], c["cos"]]).contiguous(); sin128 = torch.cat([c["sin"], c["sin"]]).contiguous()
    print("sglang form kv", kv)
    for rep in range(6):
        r = c["residual"].clone()
        ro, rr, rk, rv = ref.llama_decoder_layer_sglang(c["x"], r, c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], 1e-6, cos128, sin128)
        torch.cuda.synchronize()
        stats(f"ref run {rep} out", ro, want[0]); 
        if rep == 0: stats("ref k", rk, want[2]); stats("ref v", rv, want[3]); stats("ref residual", rr, want[1])
    r = c["residual"].clone()
    o, rr, k, v = clusterfusion.llama_decoder_layer_sglang(c["x"], r, c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], 1e-6, cos128, sin128)
    stats("ours out", o, want[0]); stats("ours k", k, want[2])
for kv in (1024,):
    d = O.make_inputs(S7, kv, seed=kv, layout="chat")
    want = O.chat_layer(d["x"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], d["cos"], d["sin"], n_heads=32, eps=1e-6, mode="eager")
    c = {k: v.cuda() for k, v in d.items()}
    print("chat form kv", kv)
    for rep in range(4):
        ro, rk, rv = ref.llama_decoder_layer(c["x"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], c["cos"], c["sin"])
        torch.cuda.synchronize()
        stats(f"ref run {rep} out", ro, want[0])
    o, k, v = clusterfusion.llama_decoder_layer(c["x"], c["weight_qkv"], c["weight_o"], c["k_cache"], c["v_cache"], c["rms_w"], c["cos"], c["sin"])
    stats("ours out", o, want[0])

"""GPU probe (not product code): the reference's own DeepSeek-MLA kernel (oracle/_ref, recompiled unmodified for sm_100a) at the
one shape its binary supports (seq_len 4096), launched a few times on identical inputs.  Prints the launch-to-launch spread and the
distance to oracle/deepseek_oracle.py.  Run it plainly and under compute-sanitizer (racecheck / initcheck / memcheck) to tell a
race or an uninitialised read inside the reference kernel from a mis-reading of its layouts by the oracle:
    compute-sanitizer --tool racecheck python tools/ref_deepseek_debug.py"""
import glob
import importlib.util
import sys

import torch

sys.path.insert(0, ".")
from oracle import deepseek_oracle as D          # noqa: E402

so = glob.glob("oracle/_ref/_clusterfusion_ref*.so")[0]
spec = importlib.util.spec_from_file_location("_clusterfusion_ref", so)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
n_runs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
d = D.make_inputs(4096, seed=4096, out_gain=2.4)
keys = ("x", "w_q_nope", "w_q_pe", "w_uk", "w_kv", "w_k_pe", "w_uv", "w_o", "ckv_cache", "rms_in_w", "rms_ckv_w", "cos", "sin")
c = [d[k].cuda() for k in keys]
want, _, _ = D.deepseek_layer(**d)
outs = []
for i in range(n_runs):
    o = ref.deepseek_decoder_layer(*c)
    torch.cuda.synchronize()
    outs.append(o.float().cpu().reshape(-1))
w = want.float().reshape(-1)
print("oracle |out| max", float(w.abs().max()))
for i, o in enumerate(outs):
    print(f"run {i}: finite {bool(torch.isfinite(o).all())} |out| max {float(o.abs().max()):.3f} max|ref - oracle| {float((o - w).abs().max()):.3f} "
          f"max|run - run0| {float((o - outs[0]).abs().max()):.3f} elements differing from run0 {int((o != outs[0]).sum())}")

"""GPU probe (not product code): bench.py's fused-FFN leg for the library named by CF_LIB_PATH (same-box A/B of two builds).
python tools/ffn_ab.py [hidden ffn]"""
import json, os, sys, torch
sys.path.insert(0, ".")
import bench
from clusterfusion_b200 import cabi
dev = torch.device("cuda", 0); cabi.load()
def tr(gr, n, w):
    for _ in range(w): gr.replay()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): gr.replay()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)
H = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
F = int(sys.argv[2]) if len(sys.argv) > 2 else 11008
r = bench.run_ffn(torch, cabi, dev, tr, bench.measured_peak_gbs()[0], hidden=H, ffn=F)
print(os.environ.get("CF_LIB_PATH", "in-tree").split("/")[-1], H, F, json.dumps({k: r[k] for k in ("us_per_layer", "frac_of_measured_peak") if k in r}), flush=True)

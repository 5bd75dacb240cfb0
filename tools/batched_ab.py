"""GPU probe (not product code): bench.py's batched paged leg for the library named by CF_LIB_PATH (same-box A/B of two builds)."""
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
for r in bench.run_batched_paged(torch, cabi, dev, tr, bench.measured_peak_gbs()[0], batches=(2, 4, 8, 16)):
    print(os.environ.get("CF_LIB_PATH", "in-tree").split("/")[-1], json.dumps({"batch": r["batch"], "batched_us": r["batched"]["us_per_layer"], "per_request_us": r.get("per_request", {}).get("us_per_layer")}), flush=True)

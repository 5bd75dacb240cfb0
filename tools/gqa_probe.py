"""GPU probe: standalone timing of the GQA kernel for the 8B shape and the 70B shard shapes on ONE GPU (no collective),
plus cluster-16 occupancy."""
import sys, json, torch
sys.path.insert(0, ".")
import bench
from clusterfusion_b200 import cabi
dev = torch.device("cuda", 0)
cabi.load()
def timed_replays(gr, n, warm):
    for _ in range(warm): gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
peak = 6552.6
peak = bench.measured_peak_gbs()[0]
for pdl, ck in ((True, False), (True, True), (False, False)):
    print("pdl", pdl, "cluster_kernel", ck)
    for tag, shape, kvs in (("llama3-8b", (4096, 32, 8), (1024, 8192)), ("70b/2", (8192, 32, 4), (1024, 16384)),
                            ("70b/4", (8192, 16, 2), (1024, 16384)), ("70b/8", (8192, 8, 1), (1024, 16384))):
        for r in bench.run_gqa_8b(torch, cabi, dev, timed_replays, peak, pdl=pdl, shape=shape, kvs=kvs, tag=tag, cluster_kernel=ck):
            print(json.dumps({k: r[k] for k in ("model", "kv_len", "us_per_layer", "achieved_gbs", "frac_of_measured_peak")}), flush=True)

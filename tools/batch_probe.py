"""GPU probe: batched paged decode (bench.run_batched_paged) only."""
import sys, json, torch
sys.path.insert(0, ".")
import bench
from clusterfusion_b200 import cabi
cabi.load()
dev = torch.device("cuda", 0)
def timed_replays(gr, n, warm):
    for _ in range(warm): gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): gr.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)
for r in bench.run_batched_paged(torch, cabi, dev, timed_replays, bench.measured_peak_gbs()[0], batches=(1, 4, 6, 8, 16)):
    print(json.dumps(r), flush=True)

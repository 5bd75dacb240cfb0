"""torchrun probe: Llama-2-70B head-parallel attention half-layer, NCCL all-reduce vs the all-reduce fused into the kernel.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29517 tools/tp_probe.py"""
import json, os, sys, torch
import torch.distributed as dist
sys.path.insert(0, ".")
import bench
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
res = bench.run_70b_sharded(torch, dist, dev, rank, world, bench.measured_peak_gbs()[0])
if rank == 0:
    for r in res:
        print(json.dumps({k: r[k] for k in ("kv_len", "world", "fused_allreduce", "us_per_layer", "achieved_gbs_per_gpu", "cuda_graph",
                                            "peer_poll_timeouts", "tokens_per_s_attn_half_80_layers")}), flush=True)
dist.barrier()
dist.destroy_process_group()

"""Two-process GPU test (needs >= 2 GPUs; skipped otherwise): the Llama-2-70B head-parallel layer -- fused kernel per
rank on its shard + ONE NCCL all-reduce -- against the full-layer oracle.  Run with `gpurun --gpus 2`."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q, fused=False):
    try:
        _worker_body(rank, world, port, q, fused)
    except BaseException as e:          # surface the failure in the parent instead of a queue timeout
        import traceback
        q.put(("error", f"rank {rank}: {e!r}\n{traceback.format_exc()}"))
        raise


def _worker_body(rank, world, port, q, fused):
    import torch.distributed as dist
    from oracle import llama_oracle as O
    from clusterfusion_b200 import sharded
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    shape = O.LayerShape(8192, 64, 8)
    kv = 300
    d = O.make_inputs(shape, kv, seed=70, layout="sglang")
    c = {k: v.cuda() for k, v in d.items()}
    sh = sharded.shard_layer(c["weight_qkv"], c["weight_o"], 64, 8, rank, world)
    lay = sharded.ShardedDecoderLayer(sh["w_qkv"], sh["w_o"], c["rms_w"], sh["n_q_heads"], sh["n_kv_heads"], 8192, 1e-5,
                                      None, world, rank=rank, fused_allreduce=fused)
    kc = sharded.shard_kv(c["k_cache"], 8, rank, world); vc = sharded.shard_kv(c["v_cache"], 8, rank, world)
    outs = []
    for it in range(7 if fused else 3):        # fused: both parities of the exchange double buffer, several times
        o, r, k, v = lay.forward(c["x"], c["residual"], kc, vc, c["cos"], c["sin"], pdl=(it == 2))
        torch.cuda.synchronize()
        outs.append(o.clone())
    if fused:
        hdr = lay.ws[:16].view(torch.int32).cpu()
        assert int(hdr[2]) == 0, "a peer poll timed out"
        assert int(hdr[3]) == 7, "one peer-stage launch counted per forward"
        assert all(torch.equal(outs[0], x_) for x_ in outs[1:]), "fused all-reduce is deterministic (rank-ordered sums)"
    ks = [torch.empty_like(k) for _ in range(world)]; vs = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(ks, k.contiguous()); dist.all_gather(vs, v.contiguous())
    allo = [torch.empty_like(o) for _ in range(world)]
    dist.all_gather(allo, o.contiguous())
    if rank == 0:
        q.put((outs[0].cpu(), outs[2].cpu(), r.cpu(), torch.cat(ks, 1).cpu(), torch.cat(vs, 1).cpu(),
               all(torch.equal(allo[0], a) for a in allo)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,fused", [(2, False), (2, True)])
def test_70b_head_parallel_nccl(world, fused):
    """fused=False: kernel + ONE NCCL all-reduce; fused=True: the all-reduce happens inside the kernel over NVLink peer
    memory (no NCCL call on the data path).  Both against the full-layer oracle; ranks must agree bit for bit."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    from oracle import llama_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, fused)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    if len(got) == 2 and isinstance(got[0], str):
        pytest.fail(got[1])
    o0, o2, r, k, v, same = got
    shape = O.LayerShape(8192, 64, 8)
    d = O.make_inputs(shape, 300, seed=70, layout="sglang")
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"],
                          1e-5, d["cos"], d["sin"], n_heads=64, n_kv_heads=8, mode="eager")
    assert same, "ranks disagree on the all-reduced output"
    assert torch.allclose(o0.float(), want[0].float(), rtol=1e-3, atol=1e-3)
    assert torch.allclose(o2.float(), want[0].float(), rtol=1e-3, atol=1e-3)
    assert torch.equal(r, want[1])
    from parity_helpers import close_k
    assert close_k(k, want[2])
    assert torch.allclose(v.view(-1).float(), want[3].view(-1).float(), rtol=1e-3, atol=1e-3)


def _stall_worker(rank, world, port, q):
    try:
        import torch.distributed as dist
        from clusterfusion_b200 import sharded
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        H, nq, nkv, kv = 8192, 32, 4, 64
        g = torch.Generator(device="cuda").manual_seed(5)
        r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device="cuda") * sc).half()
        lay = sharded.ShardedDecoderLayer(r((nq + 2 * nkv) * 128, H, sc=0.02), r(H, nq * 128, sc=0.02), (1 + 0.1 * r(H).float()).half(),
                                          nq, nkv, H, 1e-5, None, world, rank=rank, fused_allreduce=True)
        x, res, kc, vc = r(1, H), r(1, H), r(kv, nkv * 128), r(kv, nkv * 128)
        cos, sin = torch.rand(64, device="cuda"), torch.rand(64, device="cuda")
        o, _, _, _ = lay.forward(x, res, kc, vc, cos, sin)           # both ranks: a healthy launch
        torch.cuda.synchronize()
        ok0 = lay.status() == 0 and bool(torch.isfinite(o).all())
        dist.barrier()
        if rank == 0:                                                 # rank 1 "stalls": it never issues the second launch
            o, _, _, _ = lay.forward(x, res, kc, vc, cos, sin)
            torch.cuda.synchronize()
            raised = False
            try:
                lay.check()
            except RuntimeError:
                raised = True
            q.put(("ok", ok0, lay.status(), bool(torch.isnan(o.float()).any()), raised))
        dist.barrier()
        dist.destroy_process_group()
    except BaseException as e:
        import traceback
        q.put(("error", f"rank {rank}: {e!r}\n{traceback.format_exc()}"))
        raise


def test_fused_allreduce_stalled_peer_is_reported_not_silent():
    """ADVICE (round 1): a peer rank that never launches must not produce a silently wrong output.  Rank 1 skips a launch;
    rank 0's kernel polls its exchange buffer for about a second, then writes NaN and sets the workspace's sticky error word:
    ShardedDecoderLayer.status() is non-zero and .check() raises."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stall_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    if got[0] == "error":
        pytest.fail(got[1])
    _, healthy, status, has_nan, raised = got
    assert healthy, "the first (matched) launch must be clean"
    assert status != 0 and has_nan and raised

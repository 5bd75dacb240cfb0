"""World-size-2 (and 4) CPU test of the head-parallel host logic over the gloo backend: weight / KV slicing by head
(clusterfusion_b200/sharded.py) + ONE all-reduce(sum) of the fp32 O-projection partial reproduces the full
Llama-2-70B-shaped layer.  The per-rank partial is computed by the oracle here (no GPU in this suite); on the GPU
box the same slicing feeds the CUDA kernel (tests/test_gpu_parity.py::test_llama2_70b_head_parallel_shards_*)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import llama_oracle as O


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


SHAPE = O.LayerShape(2048, 16, 4)      # 70B-like ratios (4 query heads per KV head), small enough for CPU
KV = 53


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.set_num_threads(1)
        from clusterfusion_b200 import sharded          # imports the C-ABI binding lazily; no launch happens on CPU
        d = O.make_inputs(SHAPE, KV, seed=5, layout="sglang")
        sh = sharded.shard_layer(d["weight_qkv"], d["weight_o"], SHAPE.n_heads, SHAPE.n_kv_heads, rank, world)
        kc = sharded.shard_kv(d["k_cache"], SHAPE.n_kv_heads, rank, world)
        vc = sharded.shard_kv(d["v_cache"], SHAPE.n_kv_heads, rank, world)
        nq, nkv = sh["n_q_heads"], sh["n_kv_heads"]
        local = O.LayerShape(SHAPE.hidden, nq, nkv)
        wq, wk, wv = sh["w_qkv"].split([nq * 128, nkv * 128, nkv * 128], 0)
        h = (d["x"].float() + d["residual"].float()).half().float().reshape(-1)
        part, k, v = O._core(h, wq, wk, wv, sh["w_o"], kc, vc, d["rms_w"], 1e-5, d["cos"], d["sin"], "neox", local, "eager")
        part = part.reshape(1, -1).contiguous()
        dist.all_reduce(part, op=dist.ReduceOp.SUM)                    # the one collective of the layer
        ks = [torch.empty_like(k) for _ in range(world)]; vs = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(ks, k.contiguous()); dist.all_gather(vs, v.contiguous())
        if rank == 0:
            q.put(("ok", part.half(), torch.cat(ks, 0).half(), torch.cat(vs, 0).half()))
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        q.put(("err", repr(e)))
        raise


@pytest.mark.parametrize("world", [2, 4])
def test_head_parallel_allreduce_reproduces_full_layer(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    assert res[0] == "ok", res
    _, out, k, v = res
    d = O.make_inputs(SHAPE, KV, seed=5, layout="sglang")
    want = O.sglang_layer(d["x"], d["residual"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"],
                          1e-5, d["cos"], d["sin"], n_heads=SHAPE.n_heads, n_kv_heads=SHAPE.n_kv_heads, mode="eager")
    assert torch.allclose(out.float(), want[0].float(), rtol=1e-3, atol=1e-3)
    assert torch.equal(k.view(-1), want[2].view(-1)) and torch.equal(v.view(-1), want[3].view(-1))


def test_shard_shapes_and_errors():
    from clusterfusion_b200 import sharded
    wqkv = torch.arange((64 + 16) * 128 * 4, dtype=torch.float32).view((64 + 16) * 128, 4)
    wo = torch.arange(4 * 64 * 128, dtype=torch.float32).view(4, 64 * 128)
    for world in (1, 2, 4, 8):
        rows, cols = [], []
        for r in range(world):
            sh = sharded.shard_layer(wqkv, wo, 64, 8, r, world)
            assert sh["w_qkv"].shape == ((64 + 16) // world * 128, 4) and sh["w_o"].shape == (4, 64 // world * 128)
            rows.append(sh["w_qkv"]); cols.append(sh["w_o"])
        # every row of Wqkv and every column of Wo is owned by exactly one rank
        assert sorted(torch.cat(rows)[:, 0].tolist()) == sorted(wqkv[:, 0].tolist())
        assert torch.equal(torch.cat(cols, 1), wo)
    with pytest.raises(ValueError):
        sharded.shard_layer(wqkv, wo, 64, 8, 0, 3)
    kv = torch.arange(5 * 8 * 128).view(5, 8 * 128)
    assert torch.equal(torch.cat([sharded.shard_kv(kv, 8, r, 4) for r in range(4)], 1), kv)

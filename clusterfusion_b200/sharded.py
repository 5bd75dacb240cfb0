"""Head-parallel shard of the fused attention half-layer (Llama-2-70B config).

GPU g of N owns query heads [g*Hq/N, (g+1)*Hq/N) and KV heads [g*Hkv/N, (g+1)*Hkv/N): Wqkv is sharded by output
row, Wo by input column, the KV cache by head -- the Column->Row parallel scheme of the reference's eager model
(chat/llama/model.py:208-235).  Each rank runs the SAME fused kernel on its shard with
CF_FLAG_OUT_FP32_PARTIAL, which makes it emit the fp32 O-projection partial; exactly one all-reduce(sum) per
layer over NVLink (torch.distributed, NCCL) completes the layer, and the fp16 rounding happens after the
reduce so every rank holds the bit-identical result.  No collective is issued for N = 1.

`fused_allreduce=True` moves that one collective INTO the kernel: every rank's finalising CTAs push their fp32 partial
columns into all ranks' exchange buffers over NVLink peer memory (8-byte flag-in-data stores) and sum the N partials in
rank order, so the kernel's `out` is already the all-reduced fp16 result -- no NCCL launch, no conversion kernel
(`TpExchange` maps the buffers with CUDA IPC; the handles travel through torch.distributed once at start-up).

Host-side helpers only (slicing, buffer exchange, one collective); all arithmetic is in the CUDA kernel or in NCCL.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import cabi

HEAD_DIM = 128


def shard_heads(n_q_heads: int, n_kv_heads: int, rank: int, world: int):
    if n_kv_heads % world or n_q_heads % world:
        raise ValueError(f"world size {world} must divide both head counts ({n_q_heads}, {n_kv_heads})")
    nq, nkv = n_q_heads // world, n_kv_heads // world
    return (rank * nq, nq), (rank * nkv, nkv)


def shard_layer(weight_qkv: torch.Tensor, weight_o: torch.Tensor, n_q_heads: int, n_kv_heads: int,
                rank: int, world: int) -> Dict[str, torch.Tensor]:
    """nn.Linear-layout weights -> this rank's contiguous shard.

    weight_qkv [(Hq+2Hkv)*128, hidden] -> [(Hq/N + 2 Hkv/N)*128, hidden]   (rows = this rank's q | k | v heads)
    weight_o   [hidden, Hq*128]        -> [hidden, Hq/N*128]               (columns = this rank's heads)
    """
    (q0, nq), (k0, nkv) = shard_heads(n_q_heads, n_kv_heads, rank, world)
    qd, kvd = n_q_heads * HEAD_DIM, n_kv_heads * HEAD_DIM
    wq = weight_qkv[q0 * HEAD_DIM:(q0 + nq) * HEAD_DIM]
    wk = weight_qkv[qd + k0 * HEAD_DIM: qd + (k0 + nkv) * HEAD_DIM]
    wv = weight_qkv[qd + kvd + k0 * HEAD_DIM: qd + kvd + (k0 + nkv) * HEAD_DIM]
    return {"w_qkv": torch.cat([wq, wk, wv], 0).contiguous(),
            "w_o": weight_o[:, q0 * HEAD_DIM:(q0 + nq) * HEAD_DIM].contiguous(),
            "n_q_heads": nq, "n_kv_heads": nkv}


def shard_kv(cache: torch.Tensor, n_kv_heads: int, rank: int, world: int) -> torch.Tensor:
    """[kv_len, Hkv*128] -> this rank's heads [kv_len, Hkv/N*128] (contiguous copy)."""
    nkv = n_kv_heads // world
    return cache[:, rank * nkv * HEAD_DIM:(rank + 1) * nkv * HEAD_DIM].contiguous()


class TpExchange:
    """Exchange buffers of the fused all-reduce for ONE workspace: a cudaMalloc'ed, zeroed buffer on this rank
    (`cf_tp_exchange_bytes`), its CUDA IPC handle all-gathered over `group`, every peer's buffer mapped on this device.
    `ptrs[r]` is the device pointer (valid here) of rank r's buffer."""

    def __init__(self, hidden: int, rank: int, world: int, group=None):
        import ctypes as C
        import torch.distributed as dist
        lib = cabi.load()
        self.lib, self.rank, self.world = lib, rank, world
        self.nbytes = int(lib.cf_tp_exchange_bytes(hidden, world))
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        cabi.check(lib.cf_ipc_alloc(self.nbytes, C.byref(own), handle))
        self.own = own.value
        handles = [None] * world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.ptrs, self.opened = [], []
        for r in range(world):
            if r == rank:
                self.ptrs.append(self.own)
            else:
                q = C.c_void_p()
                cabi.check(lib.cf_ipc_open(handles[r], C.byref(q)))
                self.ptrs.append(q.value)
                self.opened.append(q.value)
        dist.barrier(group=group)          # nobody launches before every rank has mapped every buffer

    def fill(self, args: "cabi.CfLlamaArgs") -> None:
        args.tp_rank, args.tp_world = self.rank, self.world
        for r, q in enumerate(self.ptrs):
            args.tp_peer[r] = q

    def close(self) -> None:
        for q in self.opened:
            self.lib.cf_ipc_close(q)
        self.opened = []
        if self.own:
            self.lib.cf_ipc_free(self.own)
            self.own = None


class ShardedDecoderLayer:
    """One rank's view of a head-parallel layer.  Buffers are allocated once; `forward` is one kernel launch +
    (world > 1) one all-reduce + one tiny fp32->fp16 conversion, all on the current stream."""

    def __init__(self, weight_qkv_shard, weight_o_shard, rms_w, n_q_local, n_kv_local, hidden, eps,
                 group: Optional["torch.distributed.ProcessGroup"] = None, world: int = 1, rank: int = 0,
                 fused_allreduce: bool = False):
        self.wqkv, self.wo, self.rms_w = weight_qkv_shard, weight_o_shard, rms_w
        self.nq, self.nkv, self.hidden, self.eps = n_q_local, n_kv_local, hidden, eps
        self.group, self.world = group, world
        dev = weight_qkv_shard.device
        self.partial = torch.empty(1, hidden, dtype=torch.float32, device=dev)
        self.residual_out = torch.empty(1, hidden, dtype=torch.float16, device=dev)
        self.k_new = torch.empty(1, n_kv_local, HEAD_DIM, dtype=torch.float16, device=dev)
        self.v_new = torch.empty(1, n_kv_local, HEAD_DIM, dtype=torch.float16, device=dev)
        self.ws = torch.zeros(cabi.workspace_bytes(hidden, 1), dtype=torch.uint8, device=dev)
        self.out = torch.empty(1, hidden, dtype=torch.float16, device=dev)
        self.tp = TpExchange(hidden, rank, world, group) if (fused_allreduce and world > 1) else None

    def status(self) -> int:
        """Sticky error word of this layer's workspace (cf_workspace_status; synchronises the current stream): non-zero
        means an exchange poll inside a kernel timed out -- a peer rank stalled or died, or the group's CTAs were not
        co-resident -- and the outputs of that launch (NaN in the peer stage) must be discarded."""
        return cabi.workspace_status(self.ws.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def check(self) -> None:
        if self.status() != 0:
            raise RuntimeError("clusterfusion_b200.sharded: an exchange poll inside the fused kernel timed out "
                               "(stalled peer rank or CTAs not co-resident); the layer's outputs are invalid")

    def forward(self, x, residual, k_cache, v_cache, cos, sin, pdl: bool = False, fp16_out: bool = False, check: bool = False):
        """check=True reads the workspace's error word back after the launch (synchronises; debug mode).
        fp16_out (world == 1 only): let the kernel write the final fp16 result itself instead of the fp32 partial."""
        if self.tp is None and self.world == 1 and fp16_out:
            a = cabi.CfLlamaArgs(
                variant=cabi.CF_VARIANT_SGLANG, flags=(cabi.CF_FLAG_PDL if pdl else 0), hidden=self.hidden, n_q_heads=self.nq,
                n_kv_heads=self.nkv, head_dim=HEAD_DIM, batch=1, kv_len=k_cache.shape[0], eps=self.eps, x=x.data_ptr(),
                residual_in=residual.data_ptr(), residual_out=self.residual_out.data_ptr(), w_qkv=self.wqkv.data_ptr(),
                w_o=self.wo.data_ptr(), rms_w=self.rms_w.data_ptr(), out=self.out.data_ptr(),
                k_new=self.k_new.data_ptr(), v_new=self.v_new.data_ptr(), k_cache=k_cache.data_ptr(),
                v_cache=v_cache.data_ptr(), cos=cos.data_ptr(), sin=sin.data_ptr(), workspace=self.ws.data_ptr())
            cabi.launch(a, torch.cuda.current_stream().cuda_stream)
            if check:
                self.check()
            return self.out, self.residual_out, self.k_new, self.v_new
        if self.tp is not None:
            # all-reduce fused into the kernel: `out` is the rank-identical fp16 result, nothing else is launched
            a = cabi.CfLlamaArgs(
                variant=cabi.CF_VARIANT_SGLANG, flags=(cabi.CF_FLAG_PDL if pdl else 0), hidden=self.hidden, n_q_heads=self.nq,
                n_kv_heads=self.nkv, head_dim=HEAD_DIM, batch=1, kv_len=k_cache.shape[0], eps=self.eps, x=x.data_ptr(),
                residual_in=residual.data_ptr(), residual_out=self.residual_out.data_ptr(), w_qkv=self.wqkv.data_ptr(),
                w_o=self.wo.data_ptr(), rms_w=self.rms_w.data_ptr(), out=self.out.data_ptr(),
                k_new=self.k_new.data_ptr(), v_new=self.v_new.data_ptr(), k_cache=k_cache.data_ptr(),
                v_cache=v_cache.data_ptr(), cos=cos.data_ptr(), sin=sin.data_ptr(), workspace=self.ws.data_ptr())
            self.tp.fill(a)
            cabi.launch(a, torch.cuda.current_stream().cuda_stream)
            if check:
                self.check()
            return self.out, self.residual_out, self.k_new, self.v_new
        flags = cabi.CF_FLAG_OUT_FP32_PARTIAL | (cabi.CF_FLAG_PDL if pdl else 0)
        a = cabi.CfLlamaArgs(
            variant=cabi.CF_VARIANT_SGLANG, flags=flags, hidden=self.hidden, n_q_heads=self.nq, n_kv_heads=self.nkv,
            head_dim=HEAD_DIM, batch=1, kv_len=k_cache.shape[0], eps=self.eps, x=x.data_ptr(),
            residual_in=residual.data_ptr(), residual_out=self.residual_out.data_ptr(), w_qkv=self.wqkv.data_ptr(),
            w_o=self.wo.data_ptr(), rms_w=self.rms_w.data_ptr(), out=self.partial.data_ptr(),
            k_new=self.k_new.data_ptr(), v_new=self.v_new.data_ptr(), k_cache=k_cache.data_ptr(),
            v_cache=v_cache.data_ptr(), cos=cos.data_ptr(), sin=sin.data_ptr(), workspace=self.ws.data_ptr())
        cabi.launch(a, torch.cuda.current_stream().cuda_stream)
        if check:
            self.check()
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.partial, op=dist.ReduceOp.SUM, group=self.group)     # the ONE collective of the layer
        return self.partial.to(torch.float16), self.residual_out, self.k_new, self.v_new

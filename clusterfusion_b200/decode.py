"""Whole-model batch-1 decode loop around the fused attention half-layer.

A fairscale-free replica of what the reference's chat demo does per generated token
(/root/reference/chat/llama/model.py:570-609 block loop, :486-519 residual / FFN wiring, :407-448 SwiGLU FFN,
chat/llama/generation.py:234-271 token loop), arranged so that ONE CUDA graph replays the whole step:

* attention half-layer = `clusterfusion.llama_decoder_layer_batch_decode_sglang` (the 15-argument paged form):
  its kv length, RoPE position and KV-append slot are read from DEVICE tensors (indptr / indices / positions), so
  the same captured graph serves every position -- the 8/10-argument forms bake kv_len into TMA descriptors on
  the host.  K/V of the new token are appended in-pool by the kernel itself (no copy kernels), the residual
  stream follows the fused-add-norm convention (residual_out = x + residual).
* FFN half-layer: `ffn="torch"` = plain PyTorch (cuBLAS GEMV), exactly like the reference, whose FFN stays eager
  PyTorch (model.py:519); `ffn="fused"` = `llama_ffn_layer` (csrc/llama_ffn_kernel.cuh), which makes the
  decode step a chain of our own kernels (attention, FFN, attention, ...) with programmatic dependent launch between
  them.  Final norm and lm_head stay PyTorch.
* sampling: greedy argmax on the device; the next token id is fed back through a device tensor, so a replay
  needs no host interaction at all.  `step_host()` is the user-facing variant that takes / returns Python ints
  (one 8-byte H2D + one 8-byte D2H per token, like the reference's `.item()` per token).

`attn="eager"` swaps the fused op for the eager PyTorch attention the reference runs when USE_CLUSTER_FUSION is
off (model.py:376-405, with torch SDPA standing in for flashinfer's decode kernel) -- the GPU baseline the
speed-up of the fused op is quoted against.

Random-initialised weights (there is no checkpoint access in this environment); shapes = Llama-2-7B by default.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F

HEAD_DIM = 128


@dataclass
class ModelShape:
    n_layers: int = 32
    hidden: int = 4096
    n_heads: int = 32
    n_kv_heads: int = 32
    ffn: int = 11008
    vocab: int = 32000
    norm_eps: float = 1e-5
    rope_theta: float = 10000.0


LLAMA2_7B = ModelShape()
LLAMA3_8B = ModelShape(n_kv_heads=8, ffn=14336, vocab=128256, rope_theta=500000.0)


class LlamaDecodeEngine:
    def __init__(self, shape: ModelShape = LLAMA2_7B, max_seq: int = 2048, device="cuda", seed: int = 0,
                 attn: str = "fused", w_scale: float = 0.02, ffn: str = "torch"):
        import clusterfusion_b200 as cf       # raises ImportError if the native extension is missing: no fallback
        self._op = cf.llama_decoder_layer_batch_decode_sglang
        self._ffn_op = cf.llama_ffn_layer_out
        self._cf = cf
        self.shape, self.max_seq, self.attn, self.ffn_mode = shape, max_seq, attn, ffn
        self.dev = torch.device(device)
        s = shape
        g = torch.Generator(device=self.dev).manual_seed(seed)
        r = lambda *sz, sc=w_scale: (torch.randn(*sz, generator=g, device=self.dev, dtype=torch.float32) * sc).half()
        qd, kvd = s.n_heads * HEAD_DIM, s.n_kv_heads * HEAD_DIM
        self.embed = r(s.vocab, s.hidden, sc=1.0)
        self.lm_head = r(s.vocab, s.hidden)
        self.final_norm = torch.ones(s.hidden, dtype=torch.float16, device=self.dev)
        self.layers = []
        for _ in range(s.n_layers):
            self.layers.append(dict(
                w_qkv=r(qd + 2 * kvd, s.hidden), w_o=r(s.hidden, qd),
                attn_norm=torch.ones(s.hidden, dtype=torch.float16, device=self.dev),
                ffn_norm=torch.ones(s.hidden, dtype=torch.float16, device=self.dev),
                w13=r(2 * s.ffn, s.hidden), w2=r(s.hidden, s.ffn),
                k_pool=torch.zeros(max_seq, kvd, dtype=torch.float16, device=self.dev),
                v_pool=torch.zeros(max_seq, kvd, dtype=torch.float16, device=self.dev)))
        if ffn == "fused":
            for l in self.layers:          # one-time re-layout at load, like the reference's _build_cf_weights
                l["w2t"] = l.pop("w2").t().contiguous()
        self.k_ptrs = torch.tensor([l["k_pool"].data_ptr() for l in self.layers], dtype=torch.uint64).to(self.dev)
        self.v_ptrs = torch.tensor([l["v_pool"].data_ptr() for l in self.layers], dtype=torch.uint64).to(self.dev)
        # RoPE table [max_seq, 128] = [cos(64) | sin(64)]  (kernel_batch_sglang.cuh:322-323 layout)
        inv = 1.0 / (s.rope_theta ** (torch.arange(0, HEAD_DIM, 2, dtype=torch.float32) / HEAD_DIM))
        ang = torch.outer(torch.arange(max_seq, dtype=torch.float32), inv)
        self.cos_sin = torch.cat([ang.cos(), ang.sin()], dim=1).contiguous().to(self.dev)
        # paged-KV metadata with the identity page table: request 0 owns slots 0..pos, new token -> slot pos
        self.indices = torch.arange(max_seq, dtype=torch.int32, device=self.dev)
        self.indptr = torch.zeros(2, dtype=torch.int32, device=self.dev)
        self.positions = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.token = torch.zeros(1, dtype=torch.int64, device=self.dev)
        # activations (fixed buffers -> graph-capturable)
        H = s.hidden
        self.attn_out = torch.empty(1, H, dtype=torch.float16, device=self.dev)
        self.res_out = torch.empty(1, H, dtype=torch.float16, device=self.dev)
        self.zero_res = torch.zeros(1, H, dtype=torch.float16, device=self.dev)
        self.ffn_out = torch.empty(1, H, dtype=torch.float16, device=self.dev)
        self.ffn_res = torch.empty(1, H, dtype=torch.float16, device=self.dev)
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self._host_in = torch.zeros(1, dtype=torch.int64).pin_memory()
        self._host_out = torch.zeros(1, dtype=torch.int64).pin_memory()

    # ------------------------------------------------------------------------------------------
    def set_position(self, pos: int, fill_random: bool = True):
        """Pretend `pos` tokens are already in the cache (synthetic N(0,1) K/V rows if fill_random)."""
        assert 0 <= pos < self.max_seq
        if fill_random and pos > 0:
            g = torch.Generator(device=self.dev).manual_seed(1234)
            for l in self.layers:
                l["k_pool"][:pos] = torch.randn(pos, l["k_pool"].shape[1], generator=g, device=self.dev).half()
                l["v_pool"][:pos] = torch.randn(pos, l["v_pool"].shape[1], generator=g, device=self.dev).half()
        self.positions.fill_(pos)
        self.indptr[1] = pos + 1

    def _rmsnorm(self, h, w):
        hf = h.float()
        return (hf * torch.rsqrt(hf.pow(2).mean(-1, keepdim=True) + self.shape.norm_eps)).to(h.dtype) * w

    def _eager_attention(self, lid, x, residual):
        """The reference's eager decode attention on the GPU (model.py:376-405), NeoX RoPE, in-pool KV append."""
        s, l = self.shape, self.layers[lid]
        h = x + residual
        n = self._rmsnorm(h, l["attn_norm"])
        qkv = F.linear(n, l["w_qkv"])
        qd, kvd = s.n_heads * HEAD_DIM, s.n_kv_heads * HEAD_DIM
        q, k, v = qkv.split([qd, kvd, kvd], dim=-1)
        cs = self.cos_sin.index_select(0, self.positions)            # [1, 128]
        cos, sin = cs[:, :64], cs[:, 64:]

        def rope(t, nh):
            t = t.view(nh, HEAD_DIM).float()
            t1, t2 = t[:, :64], t[:, 64:]
            return torch.cat([t1 * cos - t2 * sin, t2 * cos + t1 * sin], dim=-1).half()
        q, k = rope(q, s.n_heads), rope(k, s.n_kv_heads)
        l["k_pool"].index_copy_(0, self.positions, k.view(1, kvd))
        l["v_pool"].index_copy_(0, self.positions, v.view(1, kvd))
        # attend over the whole pool with a length mask (graph-capturable: no host-side kv_len)
        K = l["k_pool"].view(self.max_seq, s.n_kv_heads, HEAD_DIM).transpose(0, 1)
        V = l["v_pool"].view(self.max_seq, s.n_kv_heads, HEAD_DIM).transpose(0, 1)
        mask = (torch.arange(self.max_seq, device=self.dev) <= self.positions).view(1, 1, self.max_seq)
        rep = s.n_heads // s.n_kv_heads
        o = F.scaled_dot_product_attention(q.view(s.n_kv_heads, rep, HEAD_DIM), K, V, attn_mask=mask)
        return F.linear(o.reshape(1, qd), l["w_o"]), h

    def _step_body(self):
        s = self.shape
        x = self.embed.index_select(0, self.token)                   # [1, hidden]
        residual = self.zero_res
        for lid, l in enumerate(self.layers):
            if self.attn == "fused":
                self._op(self.attn_out, self.res_out, x, residual, l["w_qkv"], l["w_o"], self.indptr, self.indices,
                         self.k_ptrs, self.v_ptrs, lid, l["attn_norm"], s.norm_eps, self.positions, self.cos_sin)
                a, h = self.attn_out, self.res_out
            else:
                a, h = self._eager_attention(lid, x, residual)
            if self.ffn_mode == "fused":
                # fused FFN half-layer: residual add + norm + SwiGLU + down projection in one launch (row f1)
                self._ffn_op(self.ffn_out, self.ffn_res, a, h, l["w13"], l["w2t"], l["ffn_norm"], s.norm_eps)
                x, residual = self.ffn_out, self.ffn_res
            else:
                h2 = h + a                                           # residual after attention (model.py:488-492)
                n = self._rmsnorm(h2, l["ffn_norm"])
                gu = F.linear(n, l["w13"])
                x = F.linear(F.silu(gu[:, :s.ffn]) * gu[:, s.ffn:], l["w2"])      # SwiGLU (model.py:447-448)
                residual = h2
        hf = self._rmsnorm(x + residual, self.final_norm)
        logits = F.linear(hf, self.lm_head)
        self.token.copy_(logits.argmax(dim=-1))
        self.positions.add_(1)
        self.indptr[1:].add_(1)

    @torch.no_grad()
    def capture(self, pdl: bool = True):
        """Capture one decode step into a CUDA graph (3 eager warm-up steps first, position restored after).
        With both halves fused, consecutive kernels are all ours and neither writes what the next one streams
        before its griddepcontrol.wait (layer l+1's weights / KV pool), so PDL is safe inside the step."""
        use_pdl = pdl and self.attn == "fused" and self.ffn_mode == "fused"
        self._cf.set_pdl(use_pdl)
        try:
            return self._capture()
        finally:
            self._cf.set_pdl(False)

    def _capture(self):
        pos0, ind0, tok0 = self.positions.clone(), self.indptr.clone(), self.token.clone()
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self._step_body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_body()
        self.positions.copy_(pos0); self.indptr.copy_(ind0); self.token.copy_(tok0)
        return self

    @torch.no_grad()
    def step(self):
        """One decoded token, device-resident (token id read from / written to self.token)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()

    def check_status(self) -> None:
        """Raise if an exchange poll inside one of the fused kernels ever timed out (CTAs of a group not co-resident because
        another kernel held SMs, a stalled peer rank): the tokens produced since are invalid.  Synchronises."""
        if self._cf.workspace_status() != 0:
            raise RuntimeError("clusterfusion_b200: an in-kernel exchange timed out; decode results since the last check are invalid")

    @torch.no_grad()
    def step_host(self, token_id: int, check_every: int = 64) -> int:
        """User-facing step: token id in from the host, next token id back to the host.  Every `check_every` tokens the
        workspaces' error words are read back as well (0 disables)."""
        self._host_in[0] = token_id
        self.token.copy_(self._host_in, non_blocking=True)
        self.step()
        self._host_out.copy_(self.token, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        self._host_steps = getattr(self, "_host_steps", 0) + 1
        if check_every and self._host_steps % check_every == 0:
            self.check_status()
        return int(self._host_out[0])

    def bytes_per_token(self, kv_len: int) -> int:
        s = self.shape
        qd, kvd = s.n_heads * HEAD_DIM, s.n_kv_heads * HEAD_DIM
        per_layer = 2 * ((qd + 2 * kvd) * s.hidden + s.hidden * qd + 3 * s.ffn * s.hidden) + 4 * kv_len * kvd
        return s.n_layers * per_layer + 2 * s.vocab * s.hidden

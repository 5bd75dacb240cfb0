"""CPU oracle for the fused DeepSeek-MLA decoder attention half-layer.

TEST INFRASTRUCTURE ONLY.  Nothing under ``clusterfusion_b200/`` (the product) may import this module.

PINNED (round 2) to the output of the reference's own kernel: the reference ships no test, no eager twin and no golden
vector for this op ("No DeepSeek test exists"); its only executable statement is the sm_90 kernel
``DeepSeekDecoderLayerKernel`` itself, which ``oracle/build_ref.sh`` recompiles unmodified for sm_100a into ``oracle/_ref``.
Launched plainly on B200 that kernel races on shared memory (compute-sanitizer racecheck: kernel.cuh:386, :486-491, :635,
dsm.cuh:37-42) and returns garbage; under compute-sanitizer memcheck the races do not fire and its output agrees with this
restatement to 5.6e-3 on outputs of magnitude 1.8.  ``oracle/gen_golden_deepseek_ref.py`` stores that output as
``tests/golden/deepseek_ref_kernel_seq4096.npz``; ``tests/test_oracle_deepseek.py`` checks this module against it at 1e-2
(the reference sums in fp16; seq_len 4096 is the only shape its binary supports).

What the reference kernel computes (``/root/reference/include/H100/deepseek/kernel.cuh``, shapes from
``config.h:1-8``: hidden 2048, 16 heads, nope 128, rope 64, kv_lora_rank 512, SEQ_LEN 4096; operator signature
``deepseek_kernel_dispatch.cu:4-18`` / ``include/pybind.cpp:45-59``):

    xn     = fp16(x * rsqrt(mean(x^2) + 1e-6) * w_rms_input)                     kernel.cuh:80-121
    q_nope = xn @ W_q_nope   [hidden, heads*128]  -> [heads, 128]                :123-165, all-reduce :290-298
    q_pe   = xn @ W_q_pe     [hidden, heads*64]   -> [heads, 64]                 :167-207
    ckv    = xn @ W_kv_nope  [hidden, 512]                                       :209-247
    k_pe   = xn @ W_k_pe     [hidden, 64]                                        :249-288
    q_pe, k_pe = rotate-half RoPE with cos[64] / sin[64]                         :300-320
    ckv_n  = fp16(ckv * rsqrt(mean(ckv^2) + 1e-6) * w_rms_ckv)                   :322-350
    q_lat  = q_nope[h] @ W_uk[:, h*512:(h+1)*512]   [128, heads*512]             :352-389, all-gather :391-398
    scores[h, t] = q_lat[h] . rows[t] / sqrt(192),  rows = ckv_cache[0:S-1, 0:512] ++ ckv_n   :400-517
                   (S = ckv_cache.shape[0]; the cache's last row is replaced by the current token, :469-470;
                    the tensor maps load latent columns 0..511 only, so q_pe / k_pe do NOT enter the scores)
    o_lat[h]     = softmax(scores[h]) @ rows                                     :441-517, cluster merge :519-598
    attn[h]      = o_lat[h] @ W_uv[:, h*128:(h+1)*128]   [512, heads*128]        :600-640
    out          = attn.flatten() @ W_o   [heads*128, hidden]  (no residual)     :642-697

``rope_scores=True`` adds the decoupled-RoPE term q_pe[h] . k_pe[t] to the scores (cache columns 512..575 and the
current token's k_pe), i.e. the complete MLA score; the default follows the reference kernel.

Arithmetic: ``mode="eager"`` rounds to fp16 wherever the kernel materialises an fp16 vector (xn, q_nope, q_pe, ckv,
k_pe, RoPE outputs, ckv_n, q_lat, o_lat, attn, out) and accumulates in fp32 inside each op (the reference kernel
additionally rounds every product and partial sum to fp16, which is not a property worth reproducing);
``mode="exact"`` is float64 throughout and measures how much of a difference is rounding.
"""
from __future__ import annotations

import math
from typing import Tuple

import torch

HIDDEN = 2048
N_HEADS = 16
NOPE = 128
ROPE = 64
LORA = 512
MLA = LORA + ROPE
EPS = 1e-6


def make_inputs(seq_len: int, seed: int = 0, out_gain: float = 4.0):
    """Seeded inputs in the reference operator's layouts; gains chosen so every intermediate is O(1)."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g) * sc).half()
    inv = torch.arange(0, ROPE // 2).float() / (ROPE // 2)
    ang = (seq_len - 1) * (10000.0 ** -inv)
    return dict(
        x=r(1, HIDDEN),
        w_q_nope=r(HIDDEN, N_HEADS * NOPE, sc=HIDDEN ** -0.5),
        w_q_pe=r(HIDDEN, N_HEADS * ROPE, sc=HIDDEN ** -0.5),
        w_uk=r(NOPE, N_HEADS * LORA, sc=NOPE ** -0.5),
        w_kv=r(HIDDEN, LORA, sc=HIDDEN ** -0.5),
        w_k_pe=r(HIDDEN, ROPE, sc=HIDDEN ** -0.5),
        w_uv=r(LORA, N_HEADS * NOPE, sc=out_gain * LORA ** -0.5),
        w_o=r(N_HEADS * NOPE, HIDDEN, sc=out_gain * (N_HEADS * NOPE) ** -0.5),
        ckv_cache=r(seq_len, MLA),
        rms_in_w=(torch.randn(HIDDEN, generator=g) * 0.1 + 1).half(),
        rms_ckv_w=(torch.randn(LORA, generator=g) * 0.1 + 1).half(),
        cos=torch.cat([ang.cos(), ang.cos()]).float(),
        sin=torch.cat([ang.sin(), ang.sin()]).float(),
    )


def _rope(v: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """kernel.cuh:300-320: out[i] = v[i] cos[i] - v[i+32] sin[i+32] (i < 32), v[i] cos[i] + v[i-32] sin[i-32] (i >= 32)."""
    h = ROPE // 2
    lo, hi = v[..., :h], v[..., h:]
    return torch.cat([lo * cos[:h] - hi * sin[h:ROPE], hi * cos[h:ROPE] + lo * sin[:h]], dim=-1)


def deepseek_layer(x, w_q_nope, w_q_pe, w_uk, w_kv, w_k_pe, w_uv, w_o, ckv_cache, rms_in_w, rms_ckv_w, cos, sin, *,
                   rope_scores: bool = False, mode: str = "eager") -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Returns (out [1, hidden], ckv_n [512], k_pe [64]) -- the last two are the current token's cache row."""
    assert mode in ("eager", "exact")
    wd = torch.float32 if mode == "eager" else torch.float64
    rnd = (lambda t: t.half().to(wd)) if mode == "eager" else (lambda t: t)
    f = lambda t: t.to(wd)
    S = ckv_cache.shape[0]
    n_heads = w_q_nope.shape[1] // NOPE
    xv = f(x).view(-1)
    xn = rnd(xv * torch.rsqrt((xv * xv).mean() + EPS) * f(rms_in_w))
    q_nope = rnd(xn @ f(w_q_nope)).view(n_heads, NOPE)
    q_pe = rnd(_rope(rnd(xn @ f(w_q_pe)).view(n_heads, ROPE), f(cos), f(sin)))
    ckv = rnd(xn @ f(w_kv))
    k_pe = rnd(_rope(rnd(xn @ f(w_k_pe)), f(cos), f(sin)))
    ckv_n = rnd(ckv * torch.rsqrt((ckv * ckv).mean() + EPS) * f(rms_ckv_w))
    q_lat = rnd(torch.einsum("hk,khn->hn", q_nope, f(w_uk).view(NOPE, n_heads, LORA)))
    rows = torch.cat([f(ckv_cache[: S - 1, :LORA]), ckv_n[None]], dim=0)                      # [S, 512]
    scores = q_lat @ rows.T
    if rope_scores:
        scores = scores + q_pe @ torch.cat([f(ckv_cache[: S - 1, LORA:]), k_pe[None]], dim=0).T
    p = torch.softmax(scores / math.sqrt(NOPE + ROPE), dim=-1)
    o_lat = rnd(p @ rows)                                                                     # [heads, 512]
    attn = rnd(torch.einsum("hk,khn->hn", o_lat, f(w_uv).view(LORA, n_heads, NOPE)))
    out = attn.reshape(1, -1) @ f(w_o)
    return out.half(), ckv_n.half(), k_pe.half()

#!/usr/bin/env python
"""bench.py -- Llama-2-7B bs=1 decode throughput of the fused attention half-layer path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--kv-len 1024] [--no-sweep]

One "step" = the hot path for ONE decoded token: 32 x llama_decoder_layer (32 distinct layers' weights
and KV caches, 4.8 GB touched per step at kv 1K, so nothing is served from the 126 MB L2).  Prints ONE
JSON line (see DESIGN.md section "Measurement" for every key).

  value      tokens/s, device-resident inputs, launches replayed from a CUDA graph through the C ABI
  e2e        tokens/s through the public operator (`clusterfusion.llama_decoder_layer`, the call
             chat/llama/model.py:358-367 makes) with the token's input copied from pinned host memory and
             the result copied back, every step, inside the timed region
  roofline   achieved algorithmic GB/s of the fused kernel vs the measured HBM peak AT KV 16K (the figure BASELINE.json's
             metric quotes), measured live in its own >= 0.25 s timed region; sub-keys: kv1k (the headline step's own
             region), paged (the 15-argument paged-KV form at 16K, random and sequential page tables), kv_sweep
             (1K/4K/16K/64K, contiguous and paged), reference_gpu_kernel_us (the reference's kernel recompiled for
             sm_100a on the same GPU)
  cpu_baseline  the reference's eager fp16 layer (restated, oracle/llama_oracle.py) on the host cores

`--impl reference` times that CPU eager path alone (the reference's own CPU-runnable implementation of the
path; its GPU kernels do not build for sm_100, /root/reference/setup.py:5-15) and prints the same line
shape with "impl": "reference".

N > 1 (torchrun): the 7B path is single-GPU by construction -> N independent replicas,
"scaling": "weak", no data-path collective; barrier + max-over-ranks timing via torch.distributed.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Llama-2-7B bs=1 decode tokens/s (fused attention half-layer path, 32 layers per token)"
HIDDEN, HEADS, D, LAYERS = 4096, 32, 128, 32
MIN_TIMED_MS = 250.0        # every headline timed region lasts at least this long, whatever --steps says


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes(kv_len: int) -> int:
    """SURVEY.md section 8d, Llama-2-7B, 8-arg (chat) form."""
    return (2 * 3 * HIDDEN * HIDDEN + 2 * HIDDEN * HIDDEN + 2 * 2 * kv_len * HIDDEN
            + 2 * HIDDEN + 2 * HIDDEN + 2 * 4 * D + 2 * HIDDEN + 2 * 2 * HIDDEN)


# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampling DURING the timed region (B200_PROFILING.md 'clocks' line)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for t, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
def cpu_eager_tokens_per_s(kv_len: int, layers_per_sample: int, reps: int, n_sets: int = 4):
    """The reference's eager fp16 decode layer (attention half) on the host cores."""
    import torch
    from oracle import llama_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(42)
    sets = []
    for _ in range(n_sets):
        w = lambda *s: (torch.randn(*s, generator=g) * 0.02).half()
        ck = torch.zeros(kv_len + 1, HEADS, D, dtype=torch.float16)
        cv = torch.zeros(kv_len + 1, HEADS, D, dtype=torch.float16)
        ck[:kv_len] = torch.randn(kv_len, HEADS, D, generator=g).half()
        cv[:kv_len] = torch.randn(kv_len, HEADS, D, generator=g).half()
        sets.append(dict(wq=w(HIDDEN, HIDDEN), wk=w(HIDDEN, HIDDEN), wv=w(HIDDEN, HIDDEN), wo=w(HIDDEN, HIDDEN),
                         ck=ck, cv=cv, rms=(1 + 0.1 * torch.randn(HIDDEN, generator=g)).half()))
    ang = torch.stack([O.rope_angles(p) for p in (kv_len,)])
    fc_full = torch.zeros(kv_len + 1, D // 2, dtype=torch.complex64)
    fc_full[kv_len] = torch.polar(torch.ones_like(ang[0]), ang[0])
    x = torch.randn(1, HIDDEN, generator=g).half()

    def sample():
        h = x
        t0 = time.perf_counter()
        with torch.no_grad():
            for i in range(layers_per_sample):
                s = sets[i % n_sets]
                h = O.eager_fp16_cpu_layer(h, s["wq"], s["wk"], s["wv"], s["wo"], s["ck"], s["cv"], s["rms"],
                                           fc_full, kv_len, eps=1e-6)
        return time.perf_counter() - t0

    sample()  # warm-up
    ts = [sample() for _ in range(reps)]
    return ts, cores


def run_reference(args, rank, world):
    if rank != 0:
        return
    import torch
    layers_per_step = 8
    ts, cores = cpu_eager_tokens_per_s(args.kv_len, layers_per_step, args.warmup + args.steps)
    ts = ts[args.warmup:] if len(ts) > args.warmup else ts
    step_s = sum(ts) / len(ts)
    tok_s = 1.0 / (step_s * (LAYERS / layers_per_step))
    sample = (f"each step = {layers_per_step} eager fp16 attention half-layers (1/4 token) on CPU, kv_len={args.kv_len}, "
              f"4 rotating weight sets; tokens/s = 1 / (step time x {LAYERS // layers_per_step})")
    line = {
        "impl": "reference", "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": len(ts), "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": {"workload": f"llama2-7b bs1 decode, kv_len={args.kv_len}: {LAYERS} x llama_decoder_layer per token "
                               "(RMSNorm+QKV+RoPE+flash-decode+O fused; FFN / lm_head are outside the hot-path scope)",
                   "hidden": HIDDEN, "heads": HEADS, "head_dim": D, "kv_len": args.kv_len, "layers_per_token": LAYERS,
                   "weights": "random N(0, 0.02^2) fp16", "replicas": 1,
                   "implementation": "the reference's eager PyTorch fp16 attention half-layer (chat/llama/model.py semantics, "
                                     "restated in oracle/llama_oracle.py) on the host CPU, all cores, rank 0 only: at N > 1 the "
                                     "driver's ratio compares N GPU replicas with this ONE host process"},
        "cpu_baseline": {"value": tok_s, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tok_s, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference GPU kernels refuse sm_100 (setup.py:5-15); its eager model needs fairscale/fire/flashinfer-CUDA, "
                "so the CPU path is the oracle's restatement of chat/llama/model.py (torch " + torch.__version__ + ")",
    }
    emit(line)


# ----------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import clusterfusion                     # public operator surface (raises if the extension is missing)
    from clusterfusion_b200 import cabi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL_DEBUG is left as the launcher set it: fd 1 was pointed at stderr in main(), so NCCL's banner / INFO lines
        # cannot get in front of the JSON line
        dist.init_process_group("nccl", device_id=dev)
    cabi.load()
    peak, peak_src = measured_peak_gbs()

    def make_layers(n, kv_len, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        L = []
        for _ in range(n):
            r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).half()
            L.append(dict(w_qkv=r(3 * HIDDEN, HIDDEN, sc=0.02), w_o=r(HIDDEN, HIDDEN, sc=0.02),
                          k=r(kv_len + 1, HIDDEN), v=r(kv_len + 1, HIDDEN),
                          rms=(1 + 0.1 * torch.randn(HIDDEN, generator=g, device=dev)).half(),
                          o=torch.empty(1, HIDDEN, dtype=torch.float16, device=dev),
                          kn=torch.empty(1, HEADS, D, dtype=torch.float16, device=dev),
                          vn=torch.empty(1, HEADS, D, dtype=torch.float16, device=dev)))
        return L

    def cos_sin(pos):     # RoPE table at position `pos`, pair-repeated as chat/llama/model.py:278-280 builds it
        a = float(pos) / (10000.0 ** (torch.arange(0, D, 2).float() / D))
        return (torch.repeat_interleave(a.cos(), 2).view(1, D).contiguous().to(dev),
                torch.repeat_interleave(a.sin(), 2).view(1, D).contiguous().to(dev))

    ws = torch.zeros(cabi.workspace_bytes(HIDDEN, 1), dtype=torch.uint8, device=dev)

    def launch_layer(x, lay, kv_len, cos, sin, stream):
        a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_CHAT, flags=(0 if args.no_pdl else cabi.CF_FLAG_PDL),
                             hidden=HIDDEN, n_q_heads=HEADS, n_kv_heads=HEADS,
                             head_dim=D, batch=1, kv_len=kv_len, eps=1e-6, x=x.data_ptr(),
                             w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(), rms_w=lay["rms"].data_ptr(),
                             out=lay["o"].data_ptr(), k_new=lay["kn"].data_ptr(), v_new=lay["vn"].data_ptr(),
                             k_cache=lay["k"].data_ptr(), v_cache=lay["v"].data_ptr(), cos=cos.data_ptr(),
                             sin=sin.data_ptr(), workspace=ws.data_ptr())
        cabi.launch(a, stream)

    def graph_of(layers, kv_len, x):
        """One CUDA graph = one pass over `layers` (layer l+1 consumes layer l's output buffer)."""
        cos, sin = cos_sin(kv_len)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for lay in layers[:2]:
                launch_layer(x, lay, kv_len, cos, sin, side.cuda_stream)   # first-call attribute setup outside capture
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            s = torch.cuda.current_stream().cuda_stream
            h = x
            for lay in layers:
                launch_layer(h, lay, kv_len, cos, sin, s)
                h = lay["o"]
        return gr, (cos, sin)

    def timed_replays(gr, n, warm):
        for _ in range(warm):
            gr.replay()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ------------------------------------------------------------------ main workload: 32 layers, kv_len
    kv = args.kv_len
    layers = make_layers(LAYERS, kv, seed=42 + rank)
    x_dev = torch.randn(1, HIDDEN, device=dev).half()
    gr, keep = graph_of(layers, kv, x_dev)

    # nn.Linear-layout copies of the same weights + paged-KV metadata for the 15-argument e2e leg (set-up, not timed)
    def build_paged():
        Lp = []
        for lay in layers:
            wq, wk, wv = lay["w_qkv"].view(3, HIDDEN, HIDDEN)                 # chat layout [W^T] -> nn.Linear layout
            Lp.append(dict(w_qkv=torch.cat([wq.t(), wk.t(), wv.t()], 0).contiguous(), w_o=lay["w_o"].t().contiguous(),
                           rms=lay["rms"], kpool=lay["k"], vpool=lay["v"]))
        kptrs = torch.tensor([l["kpool"].data_ptr() for l in Lp], dtype=torch.uint64).to(dev)
        vptrs = torch.tensor([l["vpool"].data_ptr() for l in Lp], dtype=torch.uint64).to(dev)
        indptr = torch.tensor([0, kv + 1], dtype=torch.int32, device=dev)
        indices = torch.arange(kv + 1, dtype=torch.int32, device=dev)
        positions = torch.tensor([kv], dtype=torch.int64, device=dev)
        inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2).float() / D))
        ang = torch.outer(torch.arange(kv + 1).float(), inv)
        cos_sin_tab = torch.cat([ang.cos(), ang.sin()], 1).contiguous().to(dev)
        return Lp, kptrs, vptrs, indptr, indices, positions, cos_sin_tab
    Lp, kptrs, vptrs, indptr, indices, positions, cos_sin_tab = build_paged()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.25)
    t_wall0 = time.time()
    profiling = os.environ.get("CF_PROFILE") == "1"      # ncu --profile-from-start off: capture the timed region only
    if profiling:
        for _ in range(3):
            gr.replay()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    # The timed region is `rounds` x K steps, rounds chosen so that it lasts >= MIN_TIMED_MS whatever --steps says (20 steps
    # of 0.8 ms would be a 16 ms sample); ms_per_step is the mean over all of them.  Every rank derives the same `rounds`
    # (the estimate is already the max over ranks).
    rounds = 1
    if not profiling:
        est = timed_replays(gr, 5, max(args.warmup, 3)) / 5
        rounds = max(1, math.ceil(MIN_TIMED_MS / (est * args.steps)))
    ms = timed_replays(gr, args.steps * rounds, 0)
    ms_per_step = ms / (args.steps * rounds)
    tok_s = world * 1e3 / ms_per_step
    us_layer = ms_per_step * 1e3 / LAYERS
    B = algorithmic_bytes(kv)
    ach = B / (us_layer * 1e-6) / 1e9

    # ------------------------------------------------------------------ e2e through the public operator
    # PDL is a documented switch of the public module (set_pdl): consecutive layers of a decoder stack satisfy its contract
    clusterfusion.set_pdl(not args.no_pdl)
    x_host = torch.randn(1, 1, HIDDEN).half().pin_memory()
    out_host = torch.empty(1, 1, HIDDEN, dtype=torch.float16).pin_memory()
    cos, sin = keep
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_e2e(step_fn, n):
        for _ in range(3):
            step_fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            step_fn()
        e1.record()
        torch.cuda.synchronize()
        t_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([t_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return world * n * 1e3 / t_ms

    # 8-argument chat form, driven as chat/llama/model.py:355-374 drives it: cache views, caller-side KV append, residual add
    kviews = [(lay["k"][:kv], lay["v"][:kv], lay["k"][kv:kv + 1].view(1, HEADS, D), lay["v"][kv:kv + 1].view(1, HEADS, D))
              for lay in layers]

    def e2e_step():
        h = x_host.to(dev, non_blocking=True)                                   # H2D: this token's input
        for lay, (kc, vc, kdst, vdst) in zip(layers, kviews):
            o, k_new, v_new = clusterfusion.llama_decoder_layer(h, lay["w_qkv"], lay["w_o"], kc, vc, lay["rms"], cos, sin)
            kdst.copy_(k_new)                                                   # caller-side KV append (model.py:371-372)
            vdst.copy_(v_new)
            h = h + o                                                           # caller-side residual (model.py:488-492)
        out_host.copy_(h, non_blocking=True)                                    # D2H: the step's result
        torch.cuda.current_stream().synchronize()

    e2e_steps = 2 if profiling else max(10, min(400, math.ceil(MIN_TIMED_MS / ms_per_step)))
    e2e_tok_s = timed_e2e(e2e_step, e2e_steps)

    # ---- e2e, 15-argument paged form (the call the reference README shows, README.md:55-75): KV append, residual
    #      add and RoPE-table lookup are inside the kernel, so the step is 32 operator calls and nothing else
    bufs = [torch.empty(1, HIDDEN, dtype=torch.float16, device=dev) for _ in range(4)]
    zero_res = torch.zeros(1, HIDDEN, dtype=torch.float16, device=dev)
    x_host2 = x_host.view(1, HIDDEN)
    x_static = torch.empty(1, HIDDEN, dtype=torch.float16, device=dev)

    def paged_layers(h):
        res = zero_res
        for li, lp in enumerate(Lp):
            o, ro = bufs[(2 * li) % 4], bufs[(2 * li + 1) % 4]
            clusterfusion.llama_decoder_layer(o, ro, h, res, lp["w_qkv"], lp["w_o"], indptr, indices, kptrs, vptrs, li,
                                              lp["rms"], 1e-6, positions, cos_sin_tab)
            h, res = o, ro
        return h

    def e2e_paged_step():
        x_static.copy_(x_host2, non_blocking=True)                              # H2D: this token's input
        h = paged_layers(x_static)
        out_host.view(1, HIDDEN).copy_(h, non_blocking=True)                    # D2H: the step's result
        torch.cuda.current_stream().synchronize()
    e2e_paged_tok_s = timed_e2e(e2e_paged_step, e2e_steps)

    # the same 32 public-operator calls captured ONCE into a CUDA graph (the operators never allocate or synchronise, so a
    # user can do what SGLang does for decode); per step: H2D of the token's input into the graph's static buffer, one
    # replay, D2H of the result
    e2e_graph_tok_s = None
    try:
        torch.cuda.synchronize()
        g_e2e = torch.cuda.CUDAGraph()
        out_host2 = out_host.view(1, HIDDEN)
        with torch.cuda.graph(g_e2e):
            # the step's H2D (pinned host input -> device) and D2H (result -> pinned host) are memcpy nodes of the same graph:
            # both copies run every step, inside the timed region, with one launch instead of three
            x_static.copy_(x_host2, non_blocking=True)
            h_static_out = paged_layers(x_static)
            out_host2.copy_(h_static_out, non_blocking=True)

        def e2e_graph_step():
            g_e2e.replay()
            torch.cuda.current_stream().synchronize()
        e2e_graph_tok_s = timed_e2e(e2e_graph_step, e2e_steps)
        del g_e2e
    except Exception as e:                # noqa: BLE001
        if world > 1:
            raise
        e2e_graph_tok_s = None
        sys.stderr.write(f"e2e graph leg failed: {type(e).__name__}: {e}\n")
    if profiling:
        torch.cuda.profiler.stop()
    del Lp
    del layers, gr, kviews
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ kv sweep (roofline report), rank 0 workload on every rank
    def paged_graph_of(Ls, kvs, table):
        """The 15-argument paged form over the same K/V buffers used as pools (kvs + 1 slots each), through the C ABI with the
        host's copy of the pool addresses (what the pybind shim passes): tiled TMA for runs of 16 consecutive slots,
        tile::gather4 otherwise."""
        Wp = []
        for lay in Ls:
            wq, wk, wv = lay["w_qkv"].view(3, HIDDEN, HIDDEN)
            Wp.append((torch.cat([wq.t(), wk.t(), wv.t()], 0).contiguous(), lay["w_o"].t().contiguous()))
        kp = torch.tensor([l["k"].data_ptr() for l in Ls], dtype=torch.uint64).to(dev)
        vp = torch.tensor([l["v"].data_ptr() for l in Ls], dtype=torch.uint64).to(dev)
        ip = torch.tensor([0, kvs + 1], dtype=torch.int32, device=dev)
        if table == "sequential":
            idx = torch.arange(kvs + 1, dtype=torch.int32, device=dev)
        else:
            idx = torch.randperm(kvs + 1, generator=torch.Generator().manual_seed(kvs)).int().to(dev)
        pos = torch.tensor([kvs], dtype=torch.int64, device=dev)
        inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2).float() / D))
        ang = (float(kvs) * inv).view(1, -1)
        tab = torch.zeros(kvs + 1, D, dtype=torch.float32, device=dev)
        tab[kvs] = torch.cat([ang.cos(), ang.sin()], 1).to(dev)
        ob = [torch.empty(1, HIDDEN, dtype=torch.float16, device=dev) for _ in range(4)]
        zr = torch.zeros(1, HIDDEN, dtype=torch.float16, device=dev)

        def launch(h, rr, li, o, ro, stream):
            a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_PAGED, flags=(0 if args.no_pdl else cabi.CF_FLAG_PDL), hidden=HIDDEN,
                                 n_q_heads=HEADS, n_kv_heads=HEADS, head_dim=D, batch=1, layer_id=li, eps=1e-6, x=h.data_ptr(),
                                 residual_in=rr.data_ptr(), residual_out=ro.data_ptr(), w_qkv=Wp[li][0].data_ptr(),
                                 w_o=Wp[li][1].data_ptr(), rms_w=Ls[li]["rms"].data_ptr(), out=o.data_ptr(),
                                 indptr=ip.data_ptr(), indices=idx.data_ptr(), k_pool_ptrs=kp.data_ptr(), v_pool_ptrs=vp.data_ptr(),
                                 positions=pos.data_ptr(), cos=tab.data_ptr(), k_cache=Ls[li]["k"].data_ptr(),
                                 v_cache=Ls[li]["v"].data_ptr(), workspace=ws.data_ptr())
            cabi.launch(a, stream)

        def chain(stream):
            h, rr = x_dev, zr
            for li in range(len(Ls)):
                o, ro = ob[(2 * li) % 4], ob[(2 * li + 1) % 4]
                launch(h, rr, li, o, ro, stream)
                h, rr = o, ro
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            chain(side.cuda_stream)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g3 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g3):
            chain(torch.cuda.current_stream().cuda_stream)
        return g3, (Wp, kp, vp, ip, idx, pos, tab, ob, zr)

    sweep = []
    if not args.no_sweep:
        for kvs in (1024, 4096, 16384, 65536):
            nsets = 8
            Ls = make_layers(nsets, kvs, seed=7)
            Bk = algorithmic_bytes(kvs)
            reps = max(20, int(MIN_TIMED_MS * 1e-3 / (nsets * Bk / (peak * 1e9))))      # >= 0.25 s of work at roofline speed
            row = {"kv_len": kvs, "bytes": Bk, "launches": reps * nsets, "distinct_layer_sets": nsets}

            def rate(g_):
                ms2 = timed_replays(g_, reps, 5)
                us_ = ms2 * 1e3 / (reps * nsets)
                a_ = Bk / (us_ * 1e-6) / 1e9
                return {"us_per_layer": round(us_, 3), "achieved_gbs": round(a_, 1), "frac_of_measured_peak": round(a_ / peak, 4),
                        "frac_of_8tbs": round(a_ / 8000.0, 4), "timed_region_ms": round(ms2, 1)}
            g2, keep2 = graph_of(Ls, kvs, x_dev)
            row.update(rate(g2))
            del g2
            for table in ("random", "sequential"):
                g3, keep3 = paged_graph_of(Ls, kvs, table)
                row["paged_" + table] = rate(g3)
                del g3, keep3
            sweep.append(row)
            del Ls
            torch.cuda.empty_cache()
    # ------------------------------------------------------------------ BASELINE config 4: Llama-3-8B GQA (32 Q / 8 KV), kv 8K
    def guarded(fn, *fa, **fk):
        """Auxiliary legs must not cost the headline line: at one GPU a failing leg is reported under its own key.  (With
        several ranks the legs contain collectives, so an exception on one rank has to stay fatal for all.)"""
        if world > 1:
            return fn(*fa, **fk)
        try:
            return fn(*fa, **fk)
        except Exception as e:              # noqa: BLE001
            return {"error": f"{type(e).__name__}: {e}"[:400]}

    def gqa_legs():
        return run_gqa_8b(torch, cabi, dev, timed_replays, peak, pdl=not args.no_pdl, kvs=(1024, 8192))

    gqa = ffn_res = batched = batched_gqa = deepseek = None
    if not args.no_sweep:
        deepseek = guarded(run_deepseek, torch, cabi, dev, timed_replays, peak, pdl=not args.no_pdl)
        gqa = guarded(gqa_legs)
        ffn_res = guarded(run_ffn, torch, cabi, dev, timed_replays, peak, pdl=not args.no_pdl)
        batched = guarded(run_batched_paged, torch, cabi, dev, timed_replays, peak)
        batched_gqa = {"llama3_8b": guarded(run_batched_paged, torch, cabi, dev, timed_replays, peak, batches=(4, 8), HEADS=32, KV_HEADS=8),
                       "llama3_8b_kv8k": guarded(run_batched_paged, torch, cabi, dev, timed_replays, peak, kv=8192, batches=(8,), HEADS=32, KV_HEADS=8),
                       "llama2_70b_layer": guarded(run_batched_paged, torch, cabi, dev, timed_replays, peak, nl=4, batches=(8,),
                                                   HIDDEN=8192, HEADS=64, KV_HEADS=8)}
    # ------------------------------------------------------------------ whole-model decode
    full = full8b = None
    if not args.no_sweep and not args.no_full_model:
        full = guarded(run_full_model, torch, dist, dev, world, peak)
        torch.cuda.empty_cache()
        full8b = guarded(run_full_model, torch, dist, dev, world, peak, kv0=8192, n_tok=64, shape_name="llama3-8b",
                         modes=("fused_attn_fused_ffn", "eager"))
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1)
    ref_gpu = None
    if world == 1 and not args.no_sweep:
        ref_gpu = run_ref_gpu_kernel(torch, dev)

    # ------------------------------------------------------------------ Llama-2-70B layer: unsharded (every N) and head-parallel (N > 1)
    shard70 = None
    if world in (1, 2, 4, 8) and not args.no_sweep:
        shard70 = run_70b_sharded(torch, dist, dev, rank, world, peak) if world > 1 else \
            guarded(run_70b_sharded, torch, dist, dev, rank, world, peak)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ------------------------------------------------------------------ CPU baseline (rank 0, N = 1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        lps = 8
        ts, cores = cpu_eager_tokens_per_s(kv, lps, reps=3)
        step = min(ts)
        cpu = {"value": 1.0 / (step * LAYERS / lps), "unit": "tokens/s", "cores": cores, "kind": "port",
               "sample": f"{lps} eager fp16 attention half-layers (kv_len={kv}, 4 rotating weight sets), best of 3, scaled to {LAYERS} layers/token",
               "ms_per_layer": step * 1e3 / lps}

    def traffic_of(kvx):
        prof = ROOT / "profiles" / "ncu_summary.json"
        if prof.exists():
            try:
                return json.loads(prof.read_text()).get(f"traffic_bytes_kv{kvx}")
            except Exception:
                return None
        return None

    rounds_note = f"{rounds} x {args.steps} steps timed back to back ({ms:.0f} ms); ms_per_step is their mean"
    kv1k_block = {"kv_len": kv, "achieved": ach, "frac": ach / peak, "frac_of_8tbs": ach / 8000.0, "us_per_launch": us_layer,
                  "algorithmic_bytes_per_launch": B, "traffic": traffic_of(kv), "launches_timed": args.steps * rounds * LAYERS,
                  "timed_region_ms": ms}
    k16 = next((r for r in sweep if r["kv_len"] == 16384), None)
    if k16 is not None:
        # the figure BASELINE.json's metric quotes: % of the HBM roofline at kv 16K, measured live above in its own region
        roofline = {"bound": "hbm", "achieved": k16["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": k16["frac_of_measured_peak"],
                    "traffic": traffic_of(16384), "peak_source": peak_src, "kernel": "cfb::llama_decoder_layer_kernel<CHAT,4>",
                    "kv_len": 16384, "us_per_launch": k16["us_per_layer"], "algorithmic_bytes_per_launch": k16["bytes"],
                    "frac_of_8tbs": k16["frac_of_8tbs"], "launches_timed": k16["launches"], "timed_region_ms": k16["timed_region_ms"],
                    "target": "north_star: >= 0.70 of ~8 TB/s at kv 16K = <= 71.9 us per layer",
                    "kv1k": kv1k_block,
                    "paged_kv16k": {"kernel": "cfb::llama_decoder_layer_kernel<PAGED,4> (15-argument paged-KV form, page size 1)",
                                    "random_page_table": k16["paged_random"], "sequential_page_table": k16["paged_sequential"]},
                    "kv_sweep": [{"kv_len": r["kv_len"], "us": r["us_per_layer"], "frac": r["frac_of_measured_peak"],
                                  "frac_of_8tbs": r["frac_of_8tbs"], "paged_random_us": r["paged_random"]["us_per_layer"],
                                  "paged_random_frac_of_8tbs": r["paged_random"]["frac_of_8tbs"],
                                  "paged_sequential_us": r["paged_sequential"]["us_per_layer"],
                                  "paged_sequential_frac_of_8tbs": r["paged_sequential"]["frac_of_8tbs"]} for r in sweep]}
    else:
        roofline = dict({"bound": "hbm", "peak": peak, "unit": "GB/s", "peak_source": peak_src,
                         "kernel": "cfb::llama_decoder_layer_kernel<CHAT,4>"}, **kv1k_block)
    if ref_gpu is not None and "results" in ref_gpu:
        # the comparator that matters: the reference's own kernel, recompiled for sm_100a, same GPU, same shapes
        roofline["reference_gpu_kernel_us"] = {f"kv{r['kv_len']}": {"kernel": r["us_kernel"], "per_call": r["us_per_call"]}
                                               for r in ref_gpu["results"]}
        ours = {1024: next((r["us_per_layer"] for r in sweep if r["kv_len"] == 1024), us_layer),
                16384: None if k16 is None else k16["us_per_layer"]}
        roofline["speedup_vs_reference_gpu_kernel"] = {
            f"kv{r['kv_len']}": (None if not r["us_kernel"] or not ours.get(r["kv_len"]) else round(r["us_kernel"] / ours[r["kv_len"]], 2))
            for r in ref_gpu["results"]}

    # headline e2e = the call the reference README shows (README.md:55-75): one operator call per layer does the whole
    # attention half-layer including KV append and residual.  Reported both as 32 plain stream launches per step and as one
    # replay of a CUDA graph holding the same 32 public-operator calls; the 8-argument chat form, where the caller appends K/V
    # and adds the residual with three more torch ops per layer (chat/llama/model.py:358-374, :488-492), is next to them.
    e2e_best = max(e2e_paged_tok_s, e2e_graph_tok_s or 0.0)
    e2e = {"value": e2e_best, "unit": "tokens/s", "h2d_bytes_per_step": x_host.numel() * 2,
           "d2h_bytes_per_step": out_host.numel() * 2, "steps": e2e_steps,
           "mode": "cuda_graph_of_public_calls" if e2e_best == e2e_graph_tok_s else "stream_launches",
           "value_stream_launches": e2e_paged_tok_s, "value_cuda_graph_of_public_calls": e2e_graph_tok_s,
           "value_chat_form_8arg": e2e_tok_s, "frac_of_device_resident_value": e2e_best / tok_s,
           "api": "clusterfusion.llama_decoder_layer, 15-argument paged form of the reference README (KV append + residual fused "
                  "in the kernel), 32 pybind calls per token with set_pdl(True); every step copies the token's input from "
                  "pinned host memory and the result back to pinned host memory, then synchronises.  stream_launches = the 32 "
                  "calls issued from Python every step; cuda_graph_of_public_calls = the same 32 calls captured once with "
                  "torch.cuda.graph together with the step's H2D and D2H copies (memcpy nodes of the graph) and replayed.  chat_form_8arg = the 8-argument form + caller-side KV append and residual "
                  "add exactly as chat/llama/model.py:355-374 does (3 extra torch ops per layer, no graph)"}
    config = {"workload": f"llama2-7b bs1 decode, kv_len={kv}: {LAYERS} x llama_decoder_layer per token "
                          "(RMSNorm+QKV+RoPE+flash-decode+O fused; FFN / lm_head are outside the hot-path scope)",
              "hidden": HIDDEN, "heads": HEADS, "head_dim": D, "kv_len": kv, "layers_per_token": LAYERS,
              "weights": "random N(0, 0.02^2) fp16, 32 distinct layers", "l2": "inputs larger than L2: 4.8 GB touched per step",
              "replicas": world, "per_replica_value": tok_s / world, "timed_region": rounds_note,
              "launch": "CUDA graph of 32 C-ABI launches per step" + ("" if args.no_pdl else ", programmatic dependent launch between layers (CF_FLAG_PDL)"),
              "collective": "none (7B path is single-GPU by construction: N independent replicas)"}
    if shard70 is not None:
        config["llama2_70b_head_parallel"] = shard70
    line = {
        "metric": METRIC, "value": tok_s, "unit": "tokens/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
        "config": config,
        "per_layer_us": us_layer,
        "roofline": roofline,
        "kv_sweep": sweep,
        "e2e": e2e,
        "gpu_launches": args.steps * rounds * LAYERS,
        "clocks": clocks,
        "cpu_baseline": cpu,
    }
    if ref_gpu is not None:
        line["reference_gpu_kernel"] = ref_gpu
    if full is not None:
        line["full_model_decode"] = full
    if full8b is not None:
        line["full_model_decode_llama3_8b_kv8k"] = full8b
    if ffn_res is not None:
        line["fused_ffn_half_layer"] = ffn_res
    if batched is not None:
        line["batched_paged_decode"] = batched
    if batched_gqa is not None:
        line["batched_paged_decode_gqa"] = batched_gqa
    if gqa is not None:
        line["llama3_8b_gqa"] = gqa
    if deepseek is not None:
        line["deepseek_mla_half_layer"] = deepseek
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_ref_gpu_kernel(torch, dev, kvs=(1024, 16384), nsets=8):
    """Baseline leg: the reference's OWN kernel (oracle/_ref = /root/reference/include/pybind.cpp + include/H100/*
    recompiled unmodified for sm_100a by oracle/build_ref.sh) on this GPU, same shapes, 8 rotating layer sets.
    us_per_call = CUDA events around its public operator (what a reference user gets: 3 memsets, 4 tensor-map encodes,
    2 device syncs per call, llama_kernel_dispatch.cu:15-144); us_kernel = the kernel alone from CUPTI."""
    import importlib.util
    sos = sorted((ROOT / "oracle" / "_ref").glob("_clusterfusion_ref*.so"))
    if not sos or dev.index != 0:           # the reference hard-codes cuda:0 (llama_kernel_dispatch.cu:18)
        return {"unavailable": "oracle/_ref not built" if not sos else "reference kernel is cuda:0-only"}
    try:
        spec = importlib.util.spec_from_file_location("_clusterfusion_ref", sos[0])
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        out = []
        for kv in kvs:
            g = torch.Generator(device=dev).manual_seed(5)
            r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).half()
            L = [dict(w_qkv=r(3 * HIDDEN, HIDDEN, sc=0.02), w_o=r(HIDDEN, HIDDEN, sc=0.02), k=r(kv, HIDDEN), v=r(kv, HIDDEN),
                      rms=(1 + 0.1 * r(HIDDEN).float()).half()) for _ in range(nsets)]
            x = r(1, HIDDEN)
            a = float(kv) / (10000.0 ** (torch.arange(0, D, 2).float() / D))
            cos = torch.repeat_interleave(a.cos(), 2).view(1, D).contiguous().to(dev)
            sin = torch.repeat_interleave(a.sin(), 2).view(1, D).contiguous().to(dev)

            def call(l):
                return ref.llama_decoder_layer(x, l["w_qkv"], l["w_o"], l["k"], l["v"], l["rms"], cos, sin)
            for l in L:
                call(l)
            torch.cuda.synchronize()
            reps = 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                for l in L:
                    call(l)
            e1.record(); torch.cuda.synchronize()
            us_call = e0.elapsed_time(e1) * 1e3 / (reps * nsets)
            us_kernel = None
            try:
                from torch.profiler import profile, ProfilerActivity
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    for _ in range(3):
                        for l in L:
                            call(l)
                    torch.cuda.synchronize()
                ks = [e for e in prof.events() if "LlamaDecoderLayerKernel" in e.name]
                if ks:
                    us_kernel = sum(e.device_time for e in ks) / len(ks)
            except Exception:
                us_kernel = None
            Bk = algorithmic_bytes(kv)
            out.append({"kv_len": kv, "us_per_call": round(us_call, 2), "us_kernel": None if us_kernel is None else round(us_kernel, 2),
                        "achieved_gbs_kernel": None if not us_kernel else round(Bk / (us_kernel * 1e-6) / 1e9, 1)})
            del L
            torch.cuda.empty_cache()
        return {"what": "reference LlamaDecoderLayerKernel (include/H100/llama/kernel.cuh) recompiled unmodified for sm_100a",
                "results": out}
    except Exception as e:       # a baseline leg must never take the bench down
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def run_batched_paged(torch, cabi, dev, timed_replays, peak, kv=1024, nl=8, batches=(1, 4, 8), HIDDEN=HIDDEN, HEADS=HEADS, KV_HEADS=HEADS):
    """Row f3: 15-argument paged form at batch > 1 (default: Llama-2-7B shapes), kv rows per request = `kv`.  The batched kernels
    stream every weight tile once per chunk of 4 / 8 requests (MHA: llama_decoder_batch*_kernel.cuh; grouped-query shapes:
    llama_decoder_gqa_batch_kernel.cuh); CF_FLAG_PER_REQUEST launches one cluster / group set per (request, head) like the
    reference (grid 32*4*bs, llama_kernel_batch_sglang_dispatch.cu:89)."""
    g = torch.Generator(device=dev).manual_seed(31)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).half()
    out = []
    QD, KVD = HEADS * D, KV_HEADS * D
    L = [dict(w_qkv=r(QD + 2 * KVD, HIDDEN, sc=0.02), w_o=r(HIDDEN, QD, sc=0.02), rms=(1 + 0.1 * r(HIDDEN).float()).half())
         for _ in range(nl)]
    for bs in batches:
        nslots = bs * (kv + 1)
        pools = [(r(nslots, KVD), r(nslots, KVD)) for _ in range(nl)]
        kptrs = torch.tensor([pk.data_ptr() for pk, _ in pools], dtype=torch.uint64).to(dev)
        vptrs = torch.tensor([pv.data_ptr() for _, pv in pools], dtype=torch.uint64).to(dev)
        indptr = torch.arange(0, bs + 1, dtype=torch.int32, device=dev) * (kv + 1)
        indices = torch.randperm(nslots, generator=torch.Generator().manual_seed(bs)).int().to(dev)
        positions = torch.full((bs,), kv, dtype=torch.int64, device=dev)
        inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2).float() / D))
        ang = torch.outer(torch.arange(kv + 1).float(), inv)
        cos_sin = torch.cat([ang.cos(), ang.sin()], 1).contiguous().to(dev)
        ws = torch.zeros(cabi.workspace_bytes(HIDDEN, bs), dtype=torch.uint8, device=dev)
        x = r(bs, HIDDEN); res = r(bs, HIDDEN)
        bufs = [(torch.empty(bs, HIDDEN, dtype=torch.float16, device=dev), torch.empty(bs, HIDDEN, dtype=torch.float16, device=dev))
                for _ in range(nl)]
        row = {"batch": bs, "kv_len": kv}
        for name, fl in (("batched", 0), ("per_request", cabi.CF_FLAG_PER_REQUEST)):
            if bs == 1 and name != "batched":
                continue

            def launch(h, rr, li, st):
                a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_PAGED, flags=fl | cabi.CF_FLAG_PDL, hidden=HIDDEN, n_q_heads=HEADS,
                                     n_kv_heads=KV_HEADS, head_dim=D, batch=bs, layer_id=li, eps=1e-5, x=h.data_ptr(), residual_in=rr.data_ptr(),
                                     residual_out=bufs[li][1].data_ptr(), w_qkv=L[li]["w_qkv"].data_ptr(), w_o=L[li]["w_o"].data_ptr(),
                                     rms_w=L[li]["rms"].data_ptr(), out=bufs[li][0].data_ptr(), indptr=indptr.data_ptr(),
                                     indices=indices.data_ptr(), k_pool_ptrs=kptrs.data_ptr(), v_pool_ptrs=vptrs.data_ptr(),
                                     positions=positions.data_ptr(), cos=cos_sin.data_ptr(), workspace=ws.data_ptr(),
                                     k_cache=pools[li][0].data_ptr(), v_cache=pools[li][1].data_ptr())   # host copy of the pool addresses, as the pybind shim passes
                cabi.launch(a, st)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                launch(x, res, 0, side.cuda_stream)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                st = torch.cuda.current_stream().cuda_stream
                h, rr = x, res
                for li in range(nl):
                    launch(h, rr, li, st)
                    h, rr = bufs[li]
            ms = timed_replays(gr, 30, 5)
            us = ms * 1e3 / (30 * nl)
            B = 2 * (2 * QD + 2 * KVD) * HIDDEN + bs * 4 * kv * KVD     # weights once + every request's K and V
            row[name] = {"us_per_layer": round(us, 2), "tokens_per_s_32_layers": round(bs * 1e6 / (us * LAYERS), 1),
                         "achieved_gbs_weights_once": round(B / (us * 1e-6) / 1e9, 1),
                         "frac_of_measured_peak": round(B / (us * 1e-6) / 1e9 / peak, 4)}
            del gr
        if "per_request" in row:
            row["speedup_vs_per_request"] = round(row["per_request"]["us_per_layer"] / row["batched"]["us_per_layer"], 2)
        out.append(row)
        del pools, ws
        torch.cuda.empty_cache()
    return out


def run_ffn(torch, cabi, dev, timed_replays, peak, pdl=True, hidden=4096, ffn=11008, nl=8):
    """Fused FFN half-layer alone: CUDA graph of `nl` distinct layers through the C ABI (row f1)."""
    g = torch.Generator(device=dev).manual_seed(21)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).half()
    ws = torch.zeros(cabi.workspace_bytes(hidden, 1), dtype=torch.uint8, device=dev)
    L = [dict(w13=r(2 * ffn, hidden, sc=0.02), w2t=r(ffn, hidden, sc=0.02), rms=(1 + 0.1 * r(hidden).float()).half(),
              o=torch.empty(1, hidden, dtype=torch.float16, device=dev), ro=torch.empty(1, hidden, dtype=torch.float16, device=dev))
         for _ in range(nl)]
    x, res = r(1, hidden), r(1, hidden)

    def launch(h, rr, lay, st):
        a = cabi.CfFfnArgs(flags=(cabi.CF_FLAG_PDL if pdl else 0), hidden=hidden, ffn=ffn, eps=1e-5, x=h.data_ptr(),
                           residual_in=rr.data_ptr(), w_gate_up=lay["w13"].data_ptr(), w_down_t=lay["w2t"].data_ptr(),
                           rms_w=lay["rms"].data_ptr(), out=lay["o"].data_ptr(), residual_out=lay["ro"].data_ptr(),
                           workspace=ws.data_ptr())
        cabi.launch_ffn(a, st)
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        launch(x, res, L[0], side.cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        st = torch.cuda.current_stream().cuda_stream
        h, rr = x, res
        for lay in L:
            launch(h, rr, lay, st)
            h, rr = lay["o"], lay["ro"]
    B = 2 * 3 * ffn * hidden + 2 * hidden * 5
    reps = max(20, int(0.25 / (nl * B / (peak * 1e9))))
    ms = timed_replays(gr, reps, 5)
    us = ms * 1e3 / (reps * nl)
    a = B / (us * 1e-6) / 1e9
    return {"hidden": hidden, "ffn": ffn, "us_per_layer": round(us, 3), "bytes": B, "achieved_gbs": round(a, 1),
            "frac_of_measured_peak": round(a / peak, 4), "frac_of_8tbs": round(a / 8000.0, 4),
            "kernel": "cfb::llama_ffn_layer_kernel", "launches": reps * nl}


def run_deepseek(torch, cabi, dev, timed_replays, peak, pdl=True, seq_lens=(4096, 16384), nl=16):
    """DeepSeek-MLA half-layer (row f4; the reference's config.h shapes): CUDA graph of `nl` distinct layers chained through
    x, through the C ABI.  Three kernels per call."""
    g = torch.Generator(device=dev).manual_seed(31)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).half()
    ws = torch.zeros(cabi.load().cf_deepseek_workspace_bytes(), dtype=torch.uint8, device=dev)
    out = []
    for S in seq_lens:
        L = [dict(wqn=r(2048, 2048, sc=0.022), wqp=r(2048, 1024, sc=0.022), wuk=r(128, 8192, sc=0.088), wkv=r(2048, 512, sc=0.022),
                  wkp=r(2048, 64, sc=0.022), wuv=r(512, 2048, sc=0.1), wo=r(2048, 2048, sc=0.05), cache=r(S, 576),
                  r1=(1 + 0.1 * r(2048).float()).half(), r2=(1 + 0.1 * r(512).float()).half(),
                  o=torch.empty(1, 2048, dtype=torch.float16, device=dev)) for _ in range(nl)]
        x = r(1, 2048); cos = torch.rand(64, generator=g, device=dev); sin = torch.rand(64, generator=g, device=dev)

        def launch(h, lay, st):
            a = cabi.CfDeepseekArgs(flags=cabi.CF_FLAG_PDL if pdl else 0, hidden=2048, n_heads=16, seq_len=S, eps=1e-6, x=h.data_ptr(),
                                    w_q_nope=lay["wqn"].data_ptr(), w_q_pe=lay["wqp"].data_ptr(), w_uk=lay["wuk"].data_ptr(),
                                    w_kv_nope=lay["wkv"].data_ptr(), w_k_pe=lay["wkp"].data_ptr(), w_uv=lay["wuv"].data_ptr(),
                                    w_o=lay["wo"].data_ptr(), ckv_cache=lay["cache"].data_ptr(), rms_input_w=lay["r1"].data_ptr(),
                                    rms_ckv_w=lay["r2"].data_ptr(), cos=cos.data_ptr(), sin=sin.data_ptr(), out=lay["o"].data_ptr(),
                                    workspace=ws.data_ptr())
            cabi.launch_deepseek(a, st)
        s_ = torch.cuda.Stream()
        with torch.cuda.stream(s_):
            launch(x, L[0], s_.cuda_stream)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            st = torch.cuda.current_stream().cuda_stream
            h = x
            for lay in L:
                launch(h, lay, st)
                h = lay["o"]
        B = 2 * (2048 * 2048 + 2048 * 1024 + 128 * 8192 + 2048 * 512 + 2048 * 64 + 512 * 2048 + 2048 * 2048) + (S - 1) * 1152
        ms = timed_replays(gr, 40, 5)
        us = ms * 1e3 / (40 * nl)
        a = B / (us * 1e-6) / 1e9
        out.append({"seq_len": S, "us_per_layer": round(us, 3), "bytes": B, "achieved_gbs": round(a, 1),
                    "frac_of_measured_peak": round(a / peak, 4), "kernels_per_call": 3, "launches": 40 * nl * 3,
                    "shapes": "hidden 2048, 16 heads, nope 128, rope 64, kv_lora 512 (reference config.h)"})
        del gr, L
        torch.cuda.empty_cache()
    return out


def run_full_model(torch, dist, dev, world, peak, kv0=1024, n_tok=128, shape_name="llama2-7b", modes=("fused_attn_fused_ffn", "fused", "eager")):
    """Llama-2-7B-shaped whole-model decode (random init): fused attention op + torch FFN / lm_head + device-side
    greedy sampling, one CUDA graph per token (clusterfusion_b200/decode.py).  Also the same loop with eager
    PyTorch attention (the reference's USE_CLUSTER_FUSION=false path on this GPU)."""
    from clusterfusion_b200.decode import LlamaDecodeEngine, LLAMA2_7B, LLAMA3_8B
    shape = LLAMA3_8B if shape_name == "llama3-8b" else LLAMA2_7B
    out = {"model": f"{shape_name} shapes, random init, fp16", "kv_len_start": kv0, "tokens": n_tok, "replicas": world}
    for mode in modes:
        eng = LlamaDecodeEngine(shape, max_seq=kv0 + 3 * n_tok + 16, device=dev, seed=5,
                                attn="eager" if mode == "eager" else "fused",
                                ffn="fused" if mode == "fused_attn_fused_ffn" else "torch")
        eng.set_position(kv0)
        eng.capture()
        for _ in range(8):
            eng.step()
        eng.set_position(kv0, fill_random=False)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n_tok):
            eng.step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        tps = world * n_tok * 1e3 / ms
        btok = eng.bytes_per_token(kv0 + n_tok // 2)
        out[mode] = {"tokens_per_s": round(tps, 1), "ms_per_token": round(ms / n_tok, 4),
                     "achieved_gbs_per_gpu": round(btok / (ms / n_tok * 1e-3) / 1e9, 1), "bytes_per_token": btok}
        if mode == "fused_attn_fused_ffn":
            # user-facing step: token id from the host, next token id back to the host, every token
            eng.set_position(kv0, fill_random=False)
            tok = 1
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(n_tok):
                tok = eng.step_host(tok)
            dt = time.perf_counter() - t0
            out[mode]["tokens_per_s_host_step"] = round(n_tok / dt, 1)
            out[mode]["h2d_d2h_bytes_per_token"] = 16
        del eng
        torch.cuda.empty_cache()
    if shape_name == "llama2-7b" and "fused_attn_fused_ffn" in out:
        # the one end-to-end figure the reference publishes (assets/example.gif, 1x H100, Llama-2-7B chat,
        # max_seq_len 1024, wall clock incl. prefill / detokenisation / printing) -- other hardware, other timing: context only
        out["reference_published_h100_chat_tokens_per_s"] = 127.46
        out["ratio_to_reference_published_h100"] = round(out["fused_attn_fused_ffn"]["tokens_per_s_host_step"] / 127.46, 2)
    if "fused" in out and "eager" in out:
        out["speedup_fused_attention_vs_eager"] = round(out["fused"]["tokens_per_s"] / out["eager"]["tokens_per_s"], 3)
    if "fused_attn_fused_ffn" in out and "eager" in out:
        out["speedup_fused_attention_and_ffn_vs_eager"] = round(out["fused_attn_fused_ffn"]["tokens_per_s"] / out["eager"]["tokens_per_s"], 3)
    return out


def run_gqa_8b(torch, cabi, dev, timed_replays, peak, pdl=True, shape=(4096, 32, 8), kvs=(8192,), nl=8, tag="llama3-8b"):
    """Grouped-query kernel on one GPU (10-arg sglang form through the C ABI, CUDA graph of `nl` distinct layers)."""
    H, HQ, HKV = shape
    g = torch.Generator(device=dev).manual_seed(11)
    r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).half()
    ws = torch.zeros(cabi.workspace_bytes(H, 1), dtype=torch.uint8, device=dev)
    out = []
    for kv in kvs:
        L = [dict(w_qkv=r((HQ + 2 * HKV) * 128, H, sc=0.02), w_o=r(H, HQ * 128, sc=0.02), k=r(kv, HKV * 128), v=r(kv, HKV * 128),
                  rms=(1 + 0.1 * r(H).float()).half(), o=torch.empty(1, H, dtype=torch.float16, device=dev),
                  ro=torch.empty(1, H, dtype=torch.float16, device=dev), kn=torch.empty(HKV * 128, dtype=torch.float16, device=dev),
                  vn=torch.empty(HKV * 128, dtype=torch.float16, device=dev)) for _ in range(nl)]
        x = r(1, H); res = r(1, H); cos = torch.rand(64, device=dev); sin = torch.rand(64, device=dev)

        def launch(h, rr, lay, stream):
            a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_SGLANG,
                                 flags=(cabi.CF_FLAG_PDL if pdl else 0),
                                 hidden=H, n_q_heads=HQ,
                                 n_kv_heads=HKV, head_dim=128, batch=1, kv_len=kv, eps=1e-5, x=h.data_ptr(), residual_in=rr.data_ptr(),
                                 residual_out=lay["ro"].data_ptr(), w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(),
                                 rms_w=lay["rms"].data_ptr(), out=lay["o"].data_ptr(), k_new=lay["kn"].data_ptr(), v_new=lay["vn"].data_ptr(),
                                 k_cache=lay["k"].data_ptr(), v_cache=lay["v"].data_ptr(), cos=cos.data_ptr(), sin=sin.data_ptr(),
                                 workspace=ws.data_ptr())
            cabi.launch(a, stream)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            launch(x, res, L[0], side.cuda_stream)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            st = torch.cuda.current_stream().cuda_stream
            h, rr = x, res
            for lay in L:
                launch(h, rr, lay, st)
                h, rr = lay["o"], lay["ro"]
        B = 2 * (HQ + 2 * HKV) * 128 * H + 2 * HQ * 128 * H + 4 * kv * HKV * 128 + 2 * H * 5 + 4 * HKV * 128 + 512
        reps = max(20, int(0.25 / (nl * B / (peak * 1e9))))
        ms = timed_replays(gr, reps, 5)
        us = ms * 1e3 / (reps * nl)
        a = B / (us * 1e-6) / 1e9
        out.append({"model": tag, "hidden": H, "q_heads": HQ, "kv_heads": HKV, "kv_len": kv, "us_per_layer": round(us, 3),
                    "bytes": B, "achieved_gbs": round(a, 1), "frac_of_measured_peak": round(a / peak, 4),
                    "frac_of_8tbs": round(a / 8000.0, 4),
                    "kernel": "cfb::llama_decoder_layer_gqa2_kernel<SGLANG,4> (G CTAs per group, L2 exchanges)"})
        del L, gr
        torch.cuda.empty_cache()
    return out


def run_70b_sharded(torch, dist, dev, rank, world, peak):
    """BASELINE config 5: Llama-2-70B attention half-layer (hidden 8192, 64 Q / 8 KV heads).
    world == 1: the UNSHARDED layer on one GPU (the t1 of the strong-scaling figure).
    world in (2, 4, 8): every rank also times the unsharded layer on its own GPU (same weights, so t1 and tN come from one
    run), then the head-parallel shards with the all-reduce fused into the kernel and with one NCCL all-reduce per layer
    (clusterfusion_b200/sharded.py).  Before timing, every rank checks the sharded result of one layer against the
    unsharded kernel's on the same inputs (rtol = atol = 1e-3; the unsharded kernel is the one tests/test_gpu_parity.py
    checks against the oracle at these kv lengths) and that all ranks hold the bit-identical output: `parity_ok`.
    strong_scaling_efficiency = t1 / (N x tN)."""
    from clusterfusion_b200 import sharded
    H70, HQ, HKV, nl = 8192, 64, 8, 8
    res = []

    def timed_chain(step, reps):
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        gr = None
        try:
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                step()
            run = gr.replay
        except Exception:
            gr, run = None, step
        for _ in range(5):
            run()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            run()
        e1.record(); torch.cuda.synchronize()
        t_ms = e0.elapsed_time(e1)
        if world > 1:
            dist.barrier()
            t = torch.tensor([t_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ms = float(t.item())
        return t_ms * 1e3 / (reps * nl), gr is not None

    for kvs in (1024, 16384):
        g = torch.Generator(device=dev).manual_seed(1000 + kvs)          # same seed on every rank: identical full weights
        r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).half()
        full = [dict(w_qkv=r((HQ + 2 * HKV) * 128, H70, sc=0.02), w_o=r(H70, HQ * 128, sc=0.02), rms=(1 + 0.1 * r(H70).float()).half(),
                     k=r(kvs, HKV * 128), v=r(kvs, HKV * 128)) for _ in range(nl)]
        x = r(1, H70); resid = r(1, H70)
        cos = torch.rand(64, generator=g, device=dev); sin = torch.rand(64, generator=g, device=dev)
        bytes_full = 2 * (HQ + 2 * HKV) * 128 * H70 + 2 * HQ * 128 * H70 + 4 * kvs * HKV * 128
        reps = 200 if kvs <= 1024 else 80

        # ---- unsharded layer on one GPU
        uns = [sharded.ShardedDecoderLayer(f["w_qkv"], f["w_o"], f["rms"], HQ, HKV, H70, 1e-5, None, 1) for f in full]

        def step_uns():
            h, rr = x, resid
            for lay, f in zip(uns, full):
                h, rr, _, _ = lay.forward(h, rr, f["k"], f["v"], cos, sin, pdl=True, fp16_out=True)
            return h
        t1, _ = timed_chain(step_uns, reps)
        want_o, want_r, _, _ = uns[0].forward(x, resid, full[0]["k"], full[0]["v"], cos, sin, fp16_out=True)
        want_o, want_r = want_o.clone(), want_r.clone()
        row = {"kv_len": kvs, "world": world, "unsharded_1gpu_us_per_layer": round(t1, 2), "bytes_full_layer": bytes_full,
               "unsharded_achieved_gbs": round(bytes_full / (t1 * 1e-6) / 1e9, 1),
               "unsharded_frac_of_measured_peak": round(bytes_full / (t1 * 1e-6) / 1e9 / peak, 4)}
        del uns
        if world > 1:
            for fused in (True, False):
                layers = []
                for f in full:
                    sh = sharded.shard_layer(f["w_qkv"], f["w_o"], HQ, HKV, rank, world)
                    layers.append((sharded.ShardedDecoderLayer(sh["w_qkv"], sh["w_o"], f["rms"], sh["n_q_heads"], sh["n_kv_heads"], H70,
                                                               1e-5, None, world, rank=rank, fused_allreduce=fused),
                                   sharded.shard_kv(f["k"], HKV, rank, world), sharded.shard_kv(f["v"], HKV, rank, world)))
                # parity gate: one layer, sharded vs unsharded on the same inputs; every rank must hold the same bits
                o, rr, _, _ = layers[0][0].forward(x, resid, layers[0][1], layers[0][2], cos, sin)
                torch.cuda.synchronize()
                ok = bool(torch.allclose(o.float(), want_o.float(), rtol=1e-3, atol=1e-3)) and bool(torch.equal(rr, want_r))
                gathered = [torch.empty_like(o) for _ in range(world)]
                dist.all_gather(gathered, o.contiguous())
                same_bits = all(bool(torch.equal(gathered[0], t_)) for t_ in gathered)
                flag = torch.tensor([1 if (ok and same_bits) else 0], device=dev)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN)
                max_diff = torch.tensor([float((o.float() - want_o.float()).abs().max())], device=dev)
                dist.all_reduce(max_diff, op=dist.ReduceOp.MAX)

                def step():
                    h, rr_ = x, resid
                    for lay, kc, vc in layers:
                        h, rr_, _, _ = lay.forward(h, rr_, kc, vc, cos, sin, pdl=fused)
                    return h
                tN, graphed = timed_chain(step, reps)
                timeouts = sum(lay.status() for lay, _, _ in layers)
                key = "fused_allreduce" if fused else "nccl_allreduce"
                row[key] = {"us_per_layer": round(tN, 2), "strong_scaling_efficiency": round(t1 / (world * tN), 4),
                            "speedup_vs_1gpu": round(t1 / tN, 3), "parity_ok": bool(flag.item()),
                            "max_abs_diff_vs_unsharded": float(max_diff.item()), "ranks_bit_identical": same_bits,
                            "cuda_graph": graphed, "peer_poll_timeouts": timeouts,
                            "achieved_gbs_per_gpu": round(bytes_full / world / (tN * 1e-6) / 1e9, 1),
                            "collective": ("all-reduce fused into the kernel: 8-byte flag-in-data stores to every peer over NVLink, "
                                           "rank-ordered sum; no NCCL call" if fused else
                                           "1 x NCCL all_reduce(fp32[8192]) per layer + fp32->fp16 cast")}
                for lay, _, _ in layers:
                    if lay.tp is not None:
                        lay.tp.close()
                del layers
                torch.cuda.empty_cache()
        res.append(row)
        del full
        torch.cuda.empty_cache()
    return res


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """Print the ONE JSON line.  Under torchrun the process's fd 1 was redirected to stderr at start-up (NCCL and other
    libraries print banners to stdout); the line goes to the saved original stdout."""
    txt = json.dumps(line) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, txt.encode())
    else:
        sys.stdout.write(txt)
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # stdout must carry exactly one JSON line: keep the real stdout aside and send everything else to stderr
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--kv-len", type=int, default=1024)
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-full-model", action="store_true")
    ap.add_argument("--no-pdl", action="store_true", help="value arm: plain stream-serialised launches instead of programmatic dependent launch")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        if args.steps > 20:
            args.steps = 20          # bounded sample: the CPU path is ~1000x slower
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()

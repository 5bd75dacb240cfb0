"""GPU probe (not product code): grouped-query paged form at batch > 1, CUDA graph of 8 distinct layers through the C ABI.
Columns: the weights-once kernel (llama_decoder_gqa_batch_kernel.cuh, default) against the group kernel launched with requests as
the slow grid dimension (CF_FLAG_PER_REQUEST).  `python tools/gqa_batch_probe.py [8b|70b]`."""
import json, sys, torch
sys.path.insert(0, ".")
from clusterfusion_b200 import cabi
dev = torch.device("cuda", 0)
cabi.load()
H, HQ, HKV, D, nl = (8192, 64, 8, 128, 4) if (len(sys.argv) > 1 and sys.argv[1] == "70b") else (4096, 32, 8, 128, 8)
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
L = [dict(w_qkv=r((HQ + 2 * HKV) * D, H, sc=0.02), w_o=r(H, HQ * D, sc=0.02), rms=(1 + 0.1 * r(H).float()).half()) for _ in range(nl)]
for kv in (1024, 8192):
    for bs in (2, 3, 4, 6, 8, 16):
        nslots = bs * (kv + 1)
        pools = [(r(nslots, HKV * D), r(nslots, HKV * D)) for _ in range(nl)]
        kptrs = torch.tensor([a.data_ptr() for a, _ in pools], dtype=torch.uint64).to(dev)
        vptrs = torch.tensor([b.data_ptr() for _, b in pools], dtype=torch.uint64).to(dev)
        indptr = torch.arange(0, bs + 1, dtype=torch.int32, device=dev) * (kv + 1)
        indices = torch.randperm(nslots).int().to(dev)
        positions = torch.full((bs,), kv, dtype=torch.int64, device=dev)
        cos_sin = torch.rand(kv + 1, D, device=dev)
        ws = torch.zeros(cabi.workspace_bytes(H, bs), dtype=torch.uint8, device=dev)
        x, res = r(bs, H), r(bs, H)
        bufs = [(torch.empty(bs, H, dtype=torch.float16, device=dev), torch.empty(bs, H, dtype=torch.float16, device=dev)) for _ in range(nl)]
        row = {"kv_len": kv, "batch": bs}
        outs = {}
        for name, fl in (("weights_once", 0), ("per_request", cabi.CF_FLAG_PER_REQUEST)):
            def launch(h, rr, li, st):
                a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_PAGED, flags=fl | cabi.CF_FLAG_PDL, hidden=H, n_q_heads=HQ, n_kv_heads=HKV, head_dim=D,
                                     batch=bs, layer_id=li, eps=1e-5, x=h.data_ptr(), residual_in=rr.data_ptr(), residual_out=bufs[li][1].data_ptr(),
                                     w_qkv=L[li]["w_qkv"].data_ptr(), w_o=L[li]["w_o"].data_ptr(), rms_w=L[li]["rms"].data_ptr(), out=bufs[li][0].data_ptr(),
                                     indptr=indptr.data_ptr(), indices=indices.data_ptr(), k_pool_ptrs=kptrs.data_ptr(), v_pool_ptrs=vptrs.data_ptr(),
                                     positions=positions.data_ptr(), cos=cos_sin.data_ptr(), workspace=ws.data_ptr(),
                                     k_cache=pools[li][0].data_ptr(), v_cache=pools[li][1].data_ptr())
                cabi.launch(a, st)
            s_ = torch.cuda.Stream()
            with torch.cuda.stream(s_):
                launch(x, res, 0, s_.cuda_stream)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st = torch.cuda.current_stream().cuda_stream
                h, rr = x, res
                for li in range(nl):
                    launch(h, rr, li, st); h, rr = bufs[li]
            for _ in range(5): g.replay()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(30): g.replay()
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / (30 * nl))
            row[name + "_us_per_layer"] = round(best, 2)
            outs[name] = bufs[-1][0].float().clone()
            del g
        row["max_abs_diff"] = float((outs["weights_once"] - outs["per_request"]).abs().max())
        row["status"] = cabi.workspace_status(ws.data_ptr())
        print(json.dumps(row), flush=True)
        del pools, ws
        torch.cuda.empty_cache()

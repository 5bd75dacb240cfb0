"""GPU probe (not product code): time the grouped-query kernel (10-argument form, CUDA graph of 8 distinct layers, PDL) for
several builds of the C-ABI library on one box.
    python tools/gqa_ab.py build/variants/libcf_r1.so clusterfusion_b200/libclusterfusion_b200.so"""
import ctypes as C, sys, torch
sys.path.insert(0, ".")
from clusterfusion_b200 import cabi
dev = torch.device("cuda", 0)
D, nl = 128, 8
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
SHAPES = (("llama3-8b", 4096, 32, 8, (1024, 8192)), ("70b full", 8192, 64, 8, (1024,)), ("70b/2", 8192, 32, 4, (1024,)),
          ("70b/4", 8192, 16, 2, (1024,)), ("70b/8", 8192, 8, 1, (1024, 16384)))
outs = {}
for path in sys.argv[1:]:
    lib = C.CDLL(path)
    lib.cf_llama_decoder_layer_launch.argtypes = [C.POINTER(cabi.CfLlamaArgs), C.c_void_p]
    lib.cf_llama_workspace_bytes.restype = C.c_size_t
    lib.cf_llama_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    for tag, H, HQ, HKV, kvs in SHAPES:
        ws = torch.zeros(lib.cf_llama_workspace_bytes(H, 1), dtype=torch.uint8, device=dev)
        for kv in kvs:
            torch.manual_seed(kv + HQ)
            L = [dict(w_qkv=r((HQ + 2 * HKV) * D, H, sc=0.02), w_o=r(H, HQ * D, sc=0.02), k=r(kv, HKV * D), v=r(kv, HKV * D), rms=r(H) * 0.1 + 1,
                      o=torch.empty(1, H, dtype=torch.float16, device=dev), ro=torch.empty(1, H, dtype=torch.float16, device=dev),
                      kn=torch.empty(HKV * D, dtype=torch.float16, device=dev), vn=torch.empty(HKV * D, dtype=torch.float16, device=dev)) for _ in range(nl)]
            x = r(1, H); res = r(1, H); cos = torch.rand(64, device=dev); sin = torch.rand(64, device=dev)

            def launch(h, rr, li, st):
                lay = L[li]
                a = cabi.CfLlamaArgs(variant=1, flags=cabi.CF_FLAG_PDL, hidden=H, n_q_heads=HQ, n_kv_heads=HKV, head_dim=D, batch=1, kv_len=kv,
                                     eps=1e-5, x=h.data_ptr(), residual_in=rr.data_ptr(), residual_out=lay["ro"].data_ptr(),
                                     w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(), rms_w=lay["rms"].data_ptr(), out=lay["o"].data_ptr(),
                                     k_new=lay["kn"].data_ptr(), v_new=lay["vn"].data_ptr(), k_cache=lay["k"].data_ptr(), v_cache=lay["v"].data_ptr(),
                                     cos=cos.data_ptr(), sin=sin.data_ptr(), workspace=ws.data_ptr())
                rc = lib.cf_llama_decoder_layer_launch(C.byref(a), C.c_void_p(st))
                assert rc == 0, rc
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                launch(x, res, 0, s.cuda_stream)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                st = torch.cuda.current_stream().cuda_stream
                h, rr = x, res
                for li in range(nl):
                    launch(h, rr, li, st); h, rr = L[li]["o"], L[li]["ro"]
            for _ in range(5): g.replay()
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(100): g.replay()
                e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) * 1e3 / (100 * nl))
            print(f"{path.split('/')[-1]:36s} {tag:10s} kv={kv:5d} us/layer={best:6.2f}", flush=True)
            key = (tag, kv); o_last = L[-1]['o'].float().clone()
            if key in outs: print('   max |diff| vs first lib:', float((outs[key] - o_last).abs().max()), 'finite', bool(torch.isfinite(o_last).all()), flush=True)
            else: outs[key] = o_last
            del L, g
            torch.cuda.empty_cache()

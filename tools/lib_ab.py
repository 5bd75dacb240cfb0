"""GPU probe: time the MHA kernel variants (chat / sglang / paged, batch 1) of two builds of the C-ABI library.
    python tools/lib_ab.py gpurun_out/libcf_premma.so clusterfusion_b200/libclusterfusion_b200.so"""
import ctypes as C, os, sys, torch
sys.path.insert(0, ".")
from clusterfusion_b200 import cabi
dev = torch.device("cuda", 0)
H, D, nl = 4096, 128, 8
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
outs = {}
for path in sys.argv[1:]:
    lib = C.CDLL(path)
    lib.cf_llama_decoder_layer_launch.argtypes = [C.POINTER(cabi.CfLlamaArgs), C.c_void_p]
    lib.cf_llama_workspace_bytes.restype = C.c_size_t
    lib.cf_llama_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
    ws = torch.zeros(lib.cf_llama_workspace_bytes(H, 1), dtype=torch.uint8, device=dev)
    for kv in (1024, 16384):
        torch.manual_seed(kv)
        L = [dict(w_qkv=r(3 * H, H, sc=0.02), w_o=r(H, H, sc=0.02), k=r(kv + 1, H), v=r(kv + 1, H), rms=r(H) * 0.1 + 1,
                  o=torch.empty(1, H, dtype=torch.float16, device=dev), ro=torch.empty(1, H, dtype=torch.float16, device=dev),
                  kn=torch.empty(H, dtype=torch.float16, device=dev), vn=torch.empty(H, dtype=torch.float16, device=dev)) for _ in range(nl)]
        x = r(1, H); res = r(1, H); cos = torch.rand(1, D, device=dev); sin = torch.rand(1, D, device=dev)
        kptrs = torch.tensor([l["k"].data_ptr() for l in L], dtype=torch.uint64).to(dev)
        vptrs = torch.tensor([l["v"].data_ptr() for l in L], dtype=torch.uint64).to(dev)
        indptr = torch.tensor([0, kv + 1], dtype=torch.int32, device=dev); indices = (torch.arange(kv + 1) if os.environ.get('CF_SEQ_PAGES') == '1' else torch.randperm(kv + 1)).int().to(dev)
        positions = torch.tensor([kv], dtype=torch.int64, device=dev); cos_sin = torch.rand(kv + 1, D, device=dev)
        for variant in (0, 1, 2):
            def launch(h, rr, li, st):
                lay = L[li]
                a = cabi.CfLlamaArgs(variant=variant, flags=cabi.CF_FLAG_PDL, layer_id=li, hidden=H, n_q_heads=32, n_kv_heads=32, head_dim=D, batch=1,
                                     kv_len=kv, eps=1e-6, x=h.data_ptr(), residual_in=rr.data_ptr(), residual_out=lay["ro"].data_ptr(),
                                     w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(), rms_w=lay["rms"].data_ptr(), out=lay["o"].data_ptr(),
                                     k_new=lay["kn"].data_ptr(), v_new=lay["vn"].data_ptr(), k_cache=lay["k"].data_ptr(), v_cache=lay["v"].data_ptr(),
                                     indptr=indptr.data_ptr(), indices=indices.data_ptr(), k_pool_ptrs=kptrs.data_ptr(), v_pool_ptrs=vptrs.data_ptr(),
                                     positions=positions.data_ptr(), cos=(cos_sin if variant == 2 else cos).data_ptr(), sin=sin.data_ptr(),
                                     workspace=ws.data_ptr())
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
            print(f"{path.split('/')[-1]:32s} kv={kv:5d} variant={('chat','sglang','paged')[variant]:6s} us/layer={best:6.2f}", flush=True)
            key = (kv, variant); o_last = L[-1]['o'].float().clone()
            if key in outs: print('   max |diff| vs first lib:', float((outs[key] - o_last).abs().max()), 'finite', bool(torch.isfinite(o_last).all()), flush=True)
            else: outs[key] = o_last
        del L

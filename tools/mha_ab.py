"""GPU probe: MHA chat kernel, flag-in-data output reduction vs red/last-arriver (default) vs CF_FLAG_LL_OUT, kv sweep."""
import sys, torch
sys.path.insert(0, ".")
from clusterfusion_b200 import cabi
cabi.load()
dev = torch.device("cuda", 0)
H, D, nl = 4096, 128, 8
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
ws = torch.zeros(cabi.workspace_bytes(H, 1), dtype=torch.uint8, device=dev)
x = r(1, H); cos = torch.rand(1, D, device=dev); sin = torch.rand(1, D, device=dev)
for kv in (1024, 4096, 16384, 65536):
    L = [dict(w_qkv=r(3 * H, H, sc=0.02), w_o=r(H, H, sc=0.02), k=r(kv, H), v=r(kv, H), rms=r(H) * 0.1 + 1,
              o=torch.empty(1, H, dtype=torch.float16, device=dev), kn=torch.empty(H, dtype=torch.float16, device=dev),
              vn=torch.empty(H, dtype=torch.float16, device=dev)) for _ in range(nl)]
    for name, fl in (("ll", cabi.CF_FLAG_PDL | cabi.CF_FLAG_LL_OUT), ("red", cabi.CF_FLAG_PDL), ("ll", cabi.CF_FLAG_PDL | cabi.CF_FLAG_LL_OUT), ("red", cabi.CF_FLAG_PDL)):
        def launch(h, lay, st):
            a = cabi.CfLlamaArgs(variant=0, flags=fl, hidden=H, n_q_heads=32, n_kv_heads=32, head_dim=D, batch=1, kv_len=kv, eps=1e-6,
                                 x=h.data_ptr(), w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(), rms_w=lay["rms"].data_ptr(),
                                 out=lay["o"].data_ptr(), k_new=lay["kn"].data_ptr(), v_new=lay["vn"].data_ptr(), k_cache=lay["k"].data_ptr(),
                                 v_cache=lay["v"].data_ptr(), cos=cos.data_ptr(), sin=sin.data_ptr(), workspace=ws.data_ptr())
            cabi.launch(a, st)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            launch(x, L[0], s.cuda_stream)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            st = torch.cuda.current_stream().cuda_stream
            h = x
            for lay in L:
                launch(h, lay, st); h = lay["o"]
        for _ in range(5): g.replay()
        torch.cuda.synchronize()
        reps = max(20, int(3000 / (kv / 1024 + 8)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): g.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"kv={kv:6d} {name:4s} us/layer={e0.elapsed_time(e1) * 1e3 / (reps * nl):7.2f}", flush=True)
    del L

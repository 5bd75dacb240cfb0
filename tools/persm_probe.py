"""GPU probe: is a CTA's pipeline HBM-bound or latency-bound?  Run the MHA chat kernel with 4/8/16/32 heads
(16..128 CTAs, identical per-CTA work) and report time per launch and per-SM throughput."""
import sys, torch
sys.path.insert(0, ".")
from clusterfusion_b200 import cabi
cabi.load()
dev = torch.device("cuda", 0)
H, D, kv, nl = 4096, 128, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 8
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
L = [dict(w_qkv=r(3 * H, H, sc=0.02), w_o=r(H, H, sc=0.02), k=r(kv, H), v=r(kv, H), rms=r(H) * 0.1 + 1,
          o=torch.empty(1, H, dtype=torch.float16, device=dev), kn=torch.empty(H, dtype=torch.float16, device=dev),
          vn=torch.empty(H, dtype=torch.float16, device=dev)) for _ in range(nl)]
x = r(1, H); cos = torch.rand(1, D, device=dev); sin = torch.rand(1, D, device=dev)
ws = torch.zeros(cabi.workspace_bytes(H, 1), dtype=torch.uint8, device=dev)
for nh in (4, 8, 16, 24, 32):
    def launch(h, lay, st):
        a = cabi.CfLlamaArgs(variant=0, flags=0, hidden=H, n_q_heads=nh, n_kv_heads=nh, head_dim=D, batch=1, kv_len=kv, eps=1e-6,
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
        for lay in L: launch(x, lay, st)
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (100 * nl)
    per_cta = (3 * 1024 * 128 * 2 + 128 * 1024 * 2 + kv // 4 * 512)
    print(f"heads={nh:2d} ctas={nh*4:3d} us/launch={us:6.2f}  per-SM GB/s={per_cta/us/1e3:6.1f}  total TB/s={per_cta*nh*4/us/1e6:5.2f}", flush=True)

"""Plain (graph-free) launch loops for ncu captures of the kernels bench.py's headline arm does not launch:
    python tools/ncu_targets.py gqa 8192     # group kernel, Llama-3-8B shapes
    python tools/ncu_targets.py ffn          # fused FFN half-layer, Llama-2-7B shapes
    python tools/ncu_targets.py deepseek 4096  # DeepSeek-MLA half-layer (3 kernels per call), through the C ABI
    python tools/ncu_targets.py paged 16384 random   # 15-argument paged-KV form, Llama-2-7B, batch 1 (random | sequential page table)
    python tools/ncu_targets.py paged 1024 random 8  # ... batch 8 (the weights-once batched kernels)
    python tools/ncu_targets.py paged 1024 random 8 8  # ... batch 8, 8 KV heads (Llama-3-8B: the grouped-query weights-once kernel)
8 distinct layer sets, 3 passes (24 launches); capture with  ncu --set full -k regex:<kernel> -s 8 -c 3 ..."""
import sys, torch
sys.path.insert(0, ".")
from clusterfusion_b200 import cabi
cabi.load()
dev = torch.device("cuda", 0)
what = sys.argv[1]
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
st = torch.cuda.current_stream().cuda_stream
if what == "gqa":
    kv = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    H, HQ, HKV = 4096, 32, 8
    ws = torch.zeros(cabi.workspace_bytes(H, 1), dtype=torch.uint8, device=dev)
    L = [dict(w_qkv=r((HQ + 2 * HKV) * 128, H, sc=0.02), w_o=r(H, HQ * 128, sc=0.02), k=r(kv, HKV * 128), v=r(kv, HKV * 128),
              rms=(1 + 0.1 * r(H).float()).half(), o=torch.empty(1, H, dtype=torch.float16, device=dev),
              ro=torch.empty(1, H, dtype=torch.float16, device=dev), kn=torch.empty(HKV * 128, dtype=torch.float16, device=dev),
              vn=torch.empty(HKV * 128, dtype=torch.float16, device=dev)) for _ in range(8)]
    x = r(1, H); res = r(1, H); cos = torch.rand(64, device=dev); sin = torch.rand(64, device=dev)
    for _ in range(3):
        for lay in L:
            a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_SGLANG, flags=0, hidden=H, n_q_heads=HQ, n_kv_heads=HKV, head_dim=128, batch=1,
                                 kv_len=kv, eps=1e-5, x=x.data_ptr(), residual_in=res.data_ptr(), residual_out=lay["ro"].data_ptr(),
                                 w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(), rms_w=lay["rms"].data_ptr(), out=lay["o"].data_ptr(),
                                 k_new=lay["kn"].data_ptr(), v_new=lay["vn"].data_ptr(), k_cache=lay["k"].data_ptr(), v_cache=lay["v"].data_ptr(),
                                 cos=cos.data_ptr(), sin=sin.data_ptr(), workspace=ws.data_ptr())
            cabi.launch(a, st)
elif what == "paged":
    kv = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    table = sys.argv[3] if len(sys.argv) > 3 else "random"
    bs = int(sys.argv[4]) if len(sys.argv) > 4 else 1
    HKV = int(sys.argv[5]) if len(sys.argv) > 5 else 32       # 8: Llama-3-8B shapes (grouped-query kernels)
    H, HQ = 4096, 32
    ws = torch.zeros(cabi.workspace_bytes(H, bs), dtype=torch.uint8, device=dev)
    L = [dict(w_qkv=r((HQ + 2 * HKV) * 128, H, sc=0.02), w_o=r(H, H, sc=0.02), k=r(bs * (kv + 1), HKV * 128), v=r(bs * (kv + 1), HKV * 128), rms=(1 + 0.1 * r(H).float()).half(),
              o=torch.empty(bs, H, dtype=torch.float16, device=dev), ro=torch.empty(bs, H, dtype=torch.float16, device=dev)) for _ in range(8)]
    kp = torch.tensor([l["k"].data_ptr() for l in L], dtype=torch.uint64).to(dev)
    vp = torch.tensor([l["v"].data_ptr() for l in L], dtype=torch.uint64).to(dev)
    indptr = (torch.arange(0, bs + 1, dtype=torch.int32) * (kv + 1)).to(dev)
    idx = (torch.arange(bs * (kv + 1)) if table == "sequential" else torch.randperm(bs * (kv + 1))).int().to(dev)
    pos = torch.full((bs,), kv, dtype=torch.int64, device=dev)
    tab = torch.rand(kv + 1, 128, device=dev)
    x = r(bs, H); res = r(bs, H)
    for _ in range(3):
        for li, lay in enumerate(L):
            a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_PAGED, flags=0, hidden=H, n_q_heads=HQ, n_kv_heads=HKV, head_dim=128, batch=bs,
                                 layer_id=li, eps=1e-5, x=x.data_ptr(), residual_in=res.data_ptr(), residual_out=lay["ro"].data_ptr(),
                                 w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(), rms_w=lay["rms"].data_ptr(), out=lay["o"].data_ptr(),
                                 indptr=indptr.data_ptr(), indices=idx.data_ptr(), k_pool_ptrs=kp.data_ptr(), v_pool_ptrs=vp.data_ptr(),
                                 positions=pos.data_ptr(), cos=tab.data_ptr(), k_cache=lay["k"].data_ptr(), v_cache=lay["v"].data_ptr(),
                                 workspace=ws.data_ptr())
            cabi.launch(a, st)
elif what == "deepseek":
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    ws = torch.zeros(cabi.load().cf_deepseek_workspace_bytes(), dtype=torch.uint8, device=dev)
    L = [dict(wqn=r(2048, 2048, sc=0.02), wqp=r(2048, 1024, sc=0.02), wuk=r(128, 8192, sc=0.09), wkv=r(2048, 512, sc=0.02),
              wkp=r(2048, 64, sc=0.02), wuv=r(512, 2048, sc=0.05), wo=r(2048, 2048, sc=0.03), cache=r(S, 576),
              r1=(1 + 0.1 * r(2048).float()).half(), r2=(1 + 0.1 * r(512).float()).half(),
              o=torch.empty(1, 2048, dtype=torch.float16, device=dev)) for _ in range(8)]
    x = r(1, 2048); cos = torch.rand(64, device=dev); sin = torch.rand(64, device=dev)
    for _ in range(3):
        for lay in L:
            a = cabi.CfDeepseekArgs(flags=0, hidden=2048, n_heads=16, seq_len=S, eps=1e-6, x=x.data_ptr(), w_q_nope=lay["wqn"].data_ptr(),
                                    w_q_pe=lay["wqp"].data_ptr(), w_uk=lay["wuk"].data_ptr(), w_kv_nope=lay["wkv"].data_ptr(),
                                    w_k_pe=lay["wkp"].data_ptr(), w_uv=lay["wuv"].data_ptr(), w_o=lay["wo"].data_ptr(),
                                    ckv_cache=lay["cache"].data_ptr(), rms_input_w=lay["r1"].data_ptr(), rms_ckv_w=lay["r2"].data_ptr(),
                                    cos=cos.data_ptr(), sin=sin.data_ptr(), out=lay["o"].data_ptr(), workspace=ws.data_ptr())
            cabi.launch_deepseek(a, st)
else:
    H, F = 4096, 11008
    ws = torch.zeros(cabi.workspace_bytes(H, 1), dtype=torch.uint8, device=dev)
    L = [dict(w13=r(2 * F, H, sc=0.02), w2t=r(F, H, sc=0.02), rms=(1 + 0.1 * r(H).float()).half(),
              o=torch.empty(1, H, dtype=torch.float16, device=dev), ro=torch.empty(1, H, dtype=torch.float16, device=dev)) for _ in range(8)]
    x, res = r(1, H), r(1, H)
    for _ in range(3):
        for lay in L:
            a = cabi.CfFfnArgs(flags=0, hidden=H, ffn=F, eps=1e-5, x=x.data_ptr(), residual_in=res.data_ptr(), w_gate_up=lay["w13"].data_ptr(),
                               w_down_t=lay["w2t"].data_ptr(), rms_w=lay["rms"].data_ptr(), out=lay["o"].data_ptr(),
                               residual_out=lay["ro"].data_ptr(), workspace=ws.data_ptr())
            cabi.launch_ffn(a, st)
torch.cuda.synchronize()
print("done", what)

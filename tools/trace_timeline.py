"""Phase timeline of the fused kernel (debug build with -DCF_TRACE, NOT the product library).
Builds a separate libclusterfusion_b200_trace.so, runs N back-to-back launches, prints per-phase statistics."""
import ctypes as C, subprocess, sys, os
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
if os.environ.get("CF_TRACE_LIB"):          # a prebuilt -DCF_TRACE library (A/B of two kernel versions on one box)
    lib_path = Path(os.environ["CF_TRACE_LIB"]).resolve()
else:
    lib_path = ROOT / "gpurun_out" / "libcf_trace.so"
    lib_path.parent.mkdir(exist_ok=True)
    extra = os.environ.get("CF_EXTRA_FLAGS", "").split()
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DCF_TRACE", *extra,
                           "-Xcompiler", "-fPIC", "-shared", "-o", str(lib_path),
                           str(ROOT / "clusterfusion_b200/csrc/llama_decoder.cu"), "-lcudart"])
from clusterfusion_b200 import cabi
lib = C.CDLL(str(lib_path))
lib.cf_llama_decoder_layer_launch.argtypes = [C.POINTER(cabi.CfLlamaArgs), C.c_void_p]
kv = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0            # 0 chat, 1 sglang
H = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
NH = int(sys.argv[5]) if len(sys.argv) > 5 else 32
NKV = int(sys.argv[6]) if len(sys.argv) > 6 else NH
BS = int(sys.argv[7]) if len(sys.argv) > 7 else 1                 # variant 2 (paged): requests per launch
D = 128
dev = "cuda"
nl = 8
def r(*s, sc=1.0): return (torch.randn(*s, device=dev) * sc).half()
layers = [dict(w_qkv=r((NH + 2 * NKV) * D, H, sc=0.02), w_o=r(H, NH * D, sc=0.02), k=r(kv + 1, NKV * D), v=r(kv + 1, NKV * D), rms=r(H) * 0.1 + 1,
               o=torch.empty(1, H, dtype=torch.float16, device=dev), ro=torch.empty(1, H, dtype=torch.float16, device=dev),
               kn=torch.empty(NH * D, dtype=torch.float16, device=dev),
               vn=torch.empty(NH * D, dtype=torch.float16, device=dev)) for _ in range(nl)]
res = r(1, H)
x = r(1, H); cos = torch.rand(1, D, device=dev); sin = torch.rand(1, D, device=dev)
lib.cf_llama_workspace_bytes.restype = C.c_size_t
lib.cf_llama_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
ws = torch.zeros(lib.cf_llama_workspace_bytes(H, BS), dtype=torch.uint8, device=dev)
if variant == 2:
    x = r(BS, H); res = r(BS, H)
    for lay in layers:
        lay["o"] = torch.empty(BS, H, dtype=torch.float16, device=dev); lay["ro"] = torch.empty(BS, H, dtype=torch.float16, device=dev)
        lay["k"] = r(BS * (kv + 1), NKV * D); lay["v"] = r(BS * (kv + 1), NKV * D)
    kptrs = torch.tensor([l["k"].data_ptr() for l in layers], dtype=torch.uint64).to(dev)
    vptrs = torch.tensor([l["v"].data_ptr() for l in layers], dtype=torch.uint64).to(dev)
    indptr = (torch.arange(0, BS + 1, dtype=torch.int32) * (kv + 1)).to(dev)
    indices = (torch.arange(BS * (kv + 1)) if os.environ.get('CF_SEQ_PAGES') == '1' else torch.randperm(BS * (kv + 1))).int().to(dev)
    positions = torch.full((BS,), kv, dtype=torch.int64, device=dev)
    cos_sin = torch.rand(kv + 1, D, device=dev)
if variant == 2 and NH == NKV:
    ncta = NH * 4 * ((((BS + 7) // 8) if BS >= 5 else ((BS + 3) // 4)) if (BS >= 2 and not (flags & 16)) else BS)
elif NH == NKV:
    ncta = NH * 4
else:
    ncl = NKV * ((NH // NKV) // 4)
    if flags & 4:                       # CF_FLAG_GQA_CLUSTER: first-generation cluster kernel
        ncta = ncl * (16 if ncl <= 4 else 8)
    elif variant == 2 and BS >= 2 and not (flags & 16) and not (BS == 2 and ncl * 16 <= 148):   # weights-once GQA kernel
        G = 64
        while G > 8 and (ncl * G > 148 or H // G < 128): G //= 2
        ncta = ncl * G * ((BS + 7) // 8)
    else:                               # group kernel: G CTAs per group, requests as the slow grid dimension
        G = 64
        while G > 8 and ncl * G * BS > 148: G //= 2
        ncta = ncl * G * BS
trace = torch.zeros(nl, ncta, 16, dtype=torch.int64, device=dev)
def launch(i, h):
    lay = layers[i]
    if variant == 2:
        a = cabi.CfLlamaArgs(variant=2, flags=flags, layer_id=i, hidden=H, n_q_heads=NH, n_kv_heads=NKV, head_dim=D, batch=BS, eps=1e-6,
                             residual_in=res.data_ptr(), residual_out=lay["ro"].data_ptr(), x=h.data_ptr(), w_qkv=lay["w_qkv"].data_ptr(),
                             w_o=lay["w_o"].data_ptr(), rms_w=lay["rms"].data_ptr(), out=lay["o"].data_ptr(), indptr=indptr.data_ptr(),
                             indices=indices.data_ptr(), k_pool_ptrs=kptrs.data_ptr(), v_pool_ptrs=vptrs.data_ptr(),
                             positions=positions.data_ptr(), cos=cos_sin.data_ptr(), workspace=ws.data_ptr(),
                             k_cache=(0 if os.environ.get('CF_NO_POOL_PTRS') else lay['k'].data_ptr()),
                             v_cache=(0 if os.environ.get('CF_NO_POOL_PTRS') else lay['v'].data_ptr()))   # host copy of the pool addresses
        rc = lib.cf_llama_decoder_layer_launch(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, rc
        return
    a = cabi.CfLlamaArgs(variant=variant, flags=flags, layer_id=i, hidden=H, n_q_heads=NH, n_kv_heads=NKV, head_dim=D, batch=1, kv_len=kv, eps=1e-6,
                         residual_in=res.data_ptr(), residual_out=lay["ro"].data_ptr(),
                         x=h.data_ptr(), w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(), rms_w=lay["rms"].data_ptr(),
                         out=lay["o"].data_ptr(), k_new=lay["kn"].data_ptr(), v_new=lay["vn"].data_ptr(),
                         k_cache=lay["k"].data_ptr(), v_cache=lay["v"].data_ptr(), cos=cos.data_ptr(), sin=sin.data_ptr(),
                         workspace=ws.data_ptr())
    rc = lib.cf_llama_decoder_layer_launch(C.byref(a), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, rc
lib.cf_debug_set_trace(C.c_void_p(trace.data_ptr()))
for rep in range(3):          # back-to-back launches; the last repetition's marks survive
    h = x
    for i in range(nl):
        launch(i, h); h = layers[i]["o"]
torch.cuda.synchronize()
t = trace.cpu().numpy().astype("int64")
if os.environ.get("CF_TRACE_SAVE"):
    import numpy as np
    np.save(os.environ["CF_TRACE_SAVE"], t)
print("variant", variant, "H", H, "heads", NH, NKV, "ctas", ncta)
names = {0: "entry", 12: "first TMA issue", 1: "rms done", 2: "qkv tiles done", 3: "xchg1 done", 4: "rope done", 5: "kv tiles done",
         6: "xchg2 done", 7: "o tiles done", 13: "prod last TMA", 8: "reds+counter", 9: "cta done", 10: "state published / 2nd-half kv done (batch8) / ag published (gqa batch)",
         11: "states read w0 / 2nd-half merges done (batch8)", 14: "states read all"}
starts = [t[i][:, 0].min() for i in range(nl)]
ends = [t[i][:, 9].max() for i in range(nl)]
print("launch-to-launch (first CTA entry) us:", [round((starts[i + 1] - starts[i]) / 1e3, 2) for i in range(nl - 1)])
print("gap prev last CTA done -> next first entry us:", [round((starts[i + 1] - ends[i]) / 1e3, 2) for i in range(nl - 1)])
for li in (nl - 2,):
    T = t[li]
    t0 = T[:, 0].min()
    print(f"layer {li}: kernel span {(T[:, 9].max() - t0) / 1e3:.2f} us (kv={kv})")
    for k in (0, 12, 1, 2, 3, 4, 5, 10, 11, 14, 6, 7, 8, 9):
        if T[:, k].max() == 0:
            continue
        v = (T[:, k] - t0) / 1e3
        print(f"  {names[k]:16s} min {v.min():7.2f}  med {sorted(v)[len(v)//2]:7.2f}  max {v.max():7.2f} us")

"""Phase timeline of the fused FFN kernel (debug build with -DCF_TRACE, NOT the product library): 8 back-to-back PDL-chained
launches over 8 distinct weight sets, per-phase statistics of the 7th.  python tools/trace_ffn.py [hidden ffn]"""
import ctypes as C, subprocess, sys, os
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
if os.environ.get("CF_TRACE_LIB"):
    lib_path = Path(os.environ["CF_TRACE_LIB"]).resolve()
else:
    lib_path = ROOT / "gpurun_out" / "libcf_trace.so"
    lib_path.parent.mkdir(exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-DCF_TRACE",
                           "-Xcompiler", "-fPIC", "-shared", "-o", str(lib_path),
                           str(ROOT / "clusterfusion_b200/csrc/llama_decoder.cu"), "-lcudart"])
from clusterfusion_b200 import cabi
lib = C.CDLL(str(lib_path))
lib.cf_llama_ffn_launch.argtypes = [C.POINTER(cabi.CfFfnArgs), C.c_void_p]
lib.cf_llama_workspace_bytes.restype = C.c_size_t
lib.cf_llama_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
H = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
F = int(sys.argv[2]) if len(sys.argv) > 2 else 11008
dev, nl = "cuda", 8
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
L = [dict(w13=r(2 * F, H, sc=0.02), w2t=r(F, H, sc=0.02), rms=(1 + 0.1 * r(H).float()).half(),
          o=torch.empty(1, H, dtype=torch.float16, device=dev), ro=torch.empty(1, H, dtype=torch.float16, device=dev)) for _ in range(nl)]
ws = torch.zeros(lib.cf_llama_workspace_bytes(H, 1), dtype=torch.uint8, device=dev)
x, res = r(1, H), r(1, H)
ncta = min(148, F // 16)
trace = torch.zeros(nl, ncta, 16, dtype=torch.int64, device=dev)
lib.cf_debug_set_trace(C.c_void_p(trace.data_ptr()))
st = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    h, rr = x, res
    for i, lay in enumerate(L):
        a = cabi.CfFfnArgs(flags=cabi.CF_FLAG_PDL | (i << 16), hidden=H, ffn=F, eps=1e-5, x=h.data_ptr(), residual_in=rr.data_ptr(),
                           w_gate_up=lay["w13"].data_ptr(), w_down_t=lay["w2t"].data_ptr(), rms_w=lay["rms"].data_ptr(),
                           out=lay["o"].data_ptr(), residual_out=lay["ro"].data_ptr(), workspace=ws.data_ptr(), workspace_batch=1)
        rc = lib.cf_llama_ffn_launch(C.byref(a), C.c_void_p(st)); assert rc == 0, rc
        h, rr = lay["o"], lay["ro"]
torch.cuda.synchronize()
t = trace.cpu().numpy().astype("int64")
names = {0: "entry", 1: "rms done", 2: "gate/up tiles done", 4: "swiglu done", 5: "down tiles done", 8: "reds+counter", 9: "cta done"}
starts = [t[i][:, 0].min() for i in range(nl)]
print("hidden", H, "ffn", F, "ctas", ncta)
print("launch-to-launch (first CTA entry) us:", [round((starts[i + 1] - starts[i]) / 1e3, 2) for i in range(nl - 1)])
T = t[nl - 2]; t0 = T[:, 0].min()
print(f"layer {nl - 2}: kernel span {(T[:, 9].max() - t0) / 1e3:.2f} us")
for k in (0, 1, 2, 4, 5, 8, 9):
    v = (T[:, k] - t0) / 1e3
    print(f"  {names[k]:20s} min {v.min():7.2f}  med {sorted(v)[len(v)//2]:7.2f}  max {v.max():7.2f} us")

"""GPU-box probe (not product code): host-side cost of one public-operator call = time to ENQUEUE n calls (n below the launch-queue
depth, so the host never waits for the device), next to the torch ops the reference's chat loop runs per layer."""
import sys, time, torch
sys.path.insert(0, ".")
import clusterfusion
dev = "cuda"
H = 4096
r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
x = r(1, H); wqkv = r(3 * H, H, sc=0.02); wo = r(H, H, sc=0.02); kc = r(64, H); vc = r(64, H); rms = r(H)
cos = torch.rand(1, 128, device=dev); sin = torch.rand(1, 128, device=dev)
kbuf = torch.zeros(128, H, dtype=torch.float16, device=dev)
kdst = kbuf[64:65].view(1, 32, 128)
def t(fn, n=400):
    for _ in range(20): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    dt = time.perf_counter() - t0
    torch.cuda.synchronize()
    return dt / n * 1e6
o, k, v = clusterfusion.llama_decoder_layer(x, wqkv, wo, kc, vc, rms, cos, sin)
print("llama_decoder_layer (8-arg) enqueue us/call:", round(t(lambda: clusterfusion.llama_decoder_layer(x, wqkv, wo, kc, vc, rms, cos, sin)), 2))
print("kdst.copy_(k_new)           enqueue us/call:", round(t(lambda: kdst.copy_(k)), 2))
print("h = h + o                   enqueue us/call:", round(t(lambda: x + o), 2))
print("slice view kc[:64]          us/call:", round(t(lambda: kbuf[:64]), 2))
out = torch.empty(1, H, dtype=torch.float16, device=dev); rout = torch.empty_like(out); res = r(1, H)
indptr = torch.tensor([0, 65], dtype=torch.int32, device=dev); indices = torch.arange(65, dtype=torch.int32, device=dev)
kp = r(80, H); vp = r(80, H)
kptrs = torch.tensor([kp.data_ptr()], dtype=torch.uint64).to(dev); vptrs = torch.tensor([vp.data_ptr()], dtype=torch.uint64).to(dev)
pos = torch.tensor([64], dtype=torch.int64, device=dev); tab = torch.rand(80, 128, device=dev)
f15 = lambda: clusterfusion.llama_decoder_layer(out, rout, x, res, wqkv, wo, indptr, indices, kptrs, vptrs, 0, rms, 1e-6, pos, tab)
f15()
print("llama_decoder_layer (15-arg) enqueue us/call:", round(t(f15), 2))

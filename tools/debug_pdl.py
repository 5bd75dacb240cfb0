import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import cabi_torch as ct
from clusterfusion_b200 import cabi
nl, kv = 6, 700
g = torch.Generator(device="cuda").manual_seed(0)
r = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device="cuda") * sc).half()
layers = [dict(w_qkv=r(3 * 4096, 4096, sc=0.02), w_o=r(4096, 4096, sc=0.02), k=r(kv, 4096), v=r(kv, 4096),
               rms=(1 + 0.1 * r(4096).float()).half()) for _ in range(nl)]
x = r(1, 4096)
cos = torch.rand(1, 128, device="cuda"); sin = torch.rand(1, 128, device="cuda")
ws = ct.workspace(4096, 1, x.device)
def chain(flags, chained=True, xs=None):
    outs, h = [], x
    for li, lay in enumerate(layers):
        if not chained and xs is not None: h = xs[li]
        o = torch.empty(1, 4096, dtype=torch.float16, device="cuda")
        kn = torch.empty(4096, dtype=torch.float16, device="cuda"); vn = torch.empty_like(kn)
        a = cabi.CfLlamaArgs(variant=0, flags=flags, hidden=4096, n_q_heads=32, n_kv_heads=32, head_dim=128, batch=1,
                             kv_len=kv, eps=1e-6, x=h.data_ptr(), w_qkv=lay["w_qkv"].data_ptr(), w_o=lay["w_o"].data_ptr(),
                             rms_w=lay["rms"].data_ptr(), out=o.data_ptr(), k_new=kn.data_ptr(), v_new=vn.data_ptr(),
                             k_cache=lay["k"].data_ptr(), v_cache=lay["v"].data_ptr(), cos=cos.data_ptr(), sin=sin.data_ptr(),
                             workspace=ws.data_ptr())
        cabi.launch(a, ct.stream_handle())
        outs += [o, kn, vn]; h = o
    torch.cuda.synchronize()
    return outs
def report(tag, A, B):
    row = []
    for i, (a, b) in enumerate(zip(A, B)):
        d = (a.float() - b.float()).abs().reshape(-1)
        bad = d > 2e-3 + 2e-3 * a.float().abs().reshape(-1)
        nb = int(bad.sum())
        if nb or float(d.max()) > 1e-3:
            idx = bad.nonzero().flatten().tolist()[:6]
            row.append(f"L{i//3}{'okv'[i%3]}:nbad={nb},max={float(d.max()):.2e},|a|max={float(a.float().abs().max()):.2f},idx={idx}")
    print(tag, row if row else "identical-within-tol", flush=True)
p0 = chain(0)
for t in range(3): report(f"plain vs plain {t}", p0, chain(0))
for t in range(5): report(f"plain vs PDL   {t}", p0, chain(cabi.CF_FLAG_PDL))
xs = [x] + [p0[3 * i] for i in range(nl - 1)]      # un-chained: every layer gets the plain chain's input
for t in range(5): report(f"unchained PDL  {t}", p0, chain(cabi.CF_FLAG_PDL, chained=False, xs=xs))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for t in range(3): report(f"side-stream PDL {t}", p0, chain(cabi.CF_FLAG_PDL))

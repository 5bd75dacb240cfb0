"""compute-sanitizer targets for the grouped-query batched kernel at the group sizes the product uses besides G = 16 (Llama-3-8B):
G = 8 (Llama-2-70B layer, 1024 columns per CTA), G = 32 / 64 (its 4- / 8-way shards; one RoPE pair per CTA and slot at G = 64).
    compute-sanitizer --tool racecheck python tools/sanitize_gqa_batch_shapes.py"""
import sys, torch
sys.path.insert(0, ".")
import clusterfusion
dev = "cuda"
for (H, HQ, HKV), lens in (((8192, 64, 8), [33, 0, 17]), ((8192, 16, 2), [40, 5, 16, 1]), ((8192, 8, 1), [31, 8, 0, 3, 20])):
    bs = len(lens)
    r = lambda *s, sc=1.0: (torch.randn(*s, device=dev) * sc).half()
    wqkv = r((HQ + 2 * HKV) * 128, H, sc=0.02); wo = r(H, HQ * 128, sc=0.02); rms = (1 + 0.1 * r(H).float()).half()
    n = sum(lens) + bs + 3
    kp = r(n, HKV * 128); vp = r(n, HKV * 128)
    kptrs = torch.tensor([kp.data_ptr()], dtype=torch.uint64).to(dev); vptrs = torch.tensor([vp.data_ptr()], dtype=torch.uint64).to(dev)
    indptr, idx, off = [0], [], 0
    perm = torch.randperm(n).tolist()
    for L in lens:
        idx += perm[off:off + L + 1]; off += L + 1; indptr.append(len(idx))
    x = r(bs, H); res = r(bs, H); out = torch.empty_like(x); ro = torch.empty_like(x)
    cs = torch.rand(64, 128, device=dev)
    clusterfusion.llama_decoder_layer_batch_decode_sglang(out, ro, x, res, wqkv, wo, torch.tensor(indptr, dtype=torch.int32, device=dev),
                                                          torch.tensor(idx, dtype=torch.int32, device=dev), kptrs, vptrs, 0, rms, 1e-5,
                                                          torch.tensor(lens, dtype=torch.int64, device=dev), cs)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    print("ok", H, HQ, HKV, bs)
print("status", clusterfusion.workspace_status())

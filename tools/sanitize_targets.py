"""One launch of every kernel at small kv_len, for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_targets.py"""
import os, sys, torch
sys.path.insert(0, ".")
ONLY = os.environ.get("CF_SANITIZE_ONLY_BATCH")      # e.g. "9": run only the paged launch with that batch (bisecting a report)
import clusterfusion
from oracle import llama_oracle as O
dev = "cuda"
S7, S8 = O.LayerShape(4096, 32, 32), O.LayerShape(4096, 32, 8)
d = {k: v.to(dev) for k, v in O.make_inputs(S7, 77, seed=1, layout="chat").items()}
clusterfusion.llama_decoder_layer(d["x"], d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], d["cos"], d["sin"])
d = {k: v.to(dev) for k, v in O.make_inputs(S7, 77, seed=2, layout="sglang").items()}
clusterfusion.llama_decoder_layer_sglang(d["x"], d["residual"].clone(), d["weight_qkv"], d["weight_o"], d["k_cache"], d["v_cache"], d["rms_w"], 1e-5, d["cos"], d["sin"])
g = {k: v.to(dev) for k, v in O.make_inputs(S8, 77, seed=3, layout="sglang").items()}
clusterfusion.llama_decoder_layer_sglang(g["x"], g["residual"].clone(), g["weight_qkv"], g["weight_o"], g["k_cache"], g["v_cache"], g["rms_w"], 1e-5, g["cos"], g["sin"])
for shape, dd, bs in ((S7, d, 1), (S7, d, 3), (S7, d, 5), (S7, d, 9), (S8, g, 2), (S8, g, 4), (S8, g, 9)):   # bs 3: chunks-of-4 kernel; bs 5: chunks-of-8 kernel;
                                                                                  # bs 9: two chunks = 256 CTAs = two waves of clusters; S8 bs 2: group kernel per request, bs 4 / 9: weights-once GQA kernel
    if ONLY and int(ONLY) != bs:
        continue
    lens = [33, 0, 17, 40, 5, 16, 1, 31, 8][:bs]
    n = sum(lens) + bs + 3
    kvd = shape.n_kv_heads * 128
    kp = torch.randn(n, kvd, device=dev).half(); vp = torch.randn(n, kvd, device=dev).half()
    kptrs = torch.tensor([kp.data_ptr()], dtype=torch.uint64).to(dev); vptrs = torch.tensor([vp.data_ptr()], dtype=torch.uint64).to(dev)
    indptr, idx, off = [0], [], 0
    perm = torch.randperm(n).tolist()
    for L in lens:
        idx += perm[off:off + L + 1]; off += L + 1; indptr.append(len(idx))
    x = torch.randn(bs, 4096, device=dev).half(); r = torch.randn(bs, 4096, device=dev).half()
    out = torch.empty_like(x); ro = torch.empty_like(x)
    cs = torch.rand(64, 128, device=dev)
    clusterfusion.llama_decoder_layer_batch_decode_sglang(out, ro, x, r, dd["weight_qkv"], dd["weight_o"], torch.tensor(indptr, dtype=torch.int32, device=dev),
                                                          torch.tensor(idx, dtype=torch.int32, device=dev), kptrs, vptrs, 0, dd["rms_w"], 1e-5,
                                                          torch.tensor(lens, dtype=torch.int64, device=dev), cs)
w13 = (torch.randn(2 * 1024, 4096, device=dev) * 0.02).half(); w2t = (torch.randn(1024, 4096, device=dev) * 0.02).half()
clusterfusion.llama_ffn_layer(d["x"], d["residual"], w13, w2t, d["rms_w"], 1e-5)
clusterfusion.rmsnorm(torch.randn(5, 4096, device=dev).half(), d["rms_w"])
from oracle import deepseek_oracle as DS
for seq_len in (70, 5000):                  # one-stage ring and the 4-stage ring of the attention kernel
    z = DS.make_inputs(seq_len, seed=4)
    clusterfusion.deepseek_decoder_layer_ex(*[z[k].to(dev) for k in ("x", "w_q_nope", "w_q_pe", "w_uk", "w_kv", "w_k_pe", "w_uv", "w_o",
                                                                     "ckv_cache", "rms_in_w", "rms_ckv_w", "cos", "sin")], True)
torch.cuda.synchronize()
print("sanitize targets done")

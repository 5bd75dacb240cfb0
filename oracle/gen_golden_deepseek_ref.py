"""Test infrastructure, NOT product code: mint a golden vector for the DeepSeek-MLA half-layer from the REFERENCE'S OWN KERNEL
(oracle/_ref = /root/reference/include/H100/deepseek/* recompiled unmodified for sm_100a by oracle/build_ref.sh).

Why it has to run under compute-sanitizer: on B200 the reference kernel has shared-memory races (racecheck: kernel.cuh:386, :486-491,
:635 and dsm.cuh:37-42 -- profiles/round2_ref_deepseek_sanitizer.txt); launched plainly it returns a different vector on every
launch, ~34 away from any sensible answer.  Under `compute-sanitizer --tool memcheck` (or initcheck) the instrumented kernel runs
slowly enough that the races do not fire, the output is the same under both tools to fp16 rounding, and it agrees with
oracle/deepseek_oracle.py to 6e-3 on outputs of magnitude 1.8.  That deterministic output is the fixture:

    gpurun -- 'compute-sanitizer --tool memcheck python oracle/gen_golden_deepseek_ref.py gpurun_out/deepseek_ref_kernel_seq4096.npz'
    cp gpurun_out/deepseek_ref_kernel_seq4096.npz tests/golden/

Inputs are regenerated from the seed by the tests (oracle/deepseek_oracle.py::make_inputs); their SHA-256 is stored next to the
output.  seq_len 4096 is the only shape the reference binary supports (include/H100/deepseek/config.h)."""
import glob
import hashlib
import importlib.util
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import deepseek_oracle as D          # noqa: E402

KEYS = ("x", "w_q_nope", "w_q_pe", "w_uk", "w_kv", "w_k_pe", "w_uv", "w_o", "ckv_cache", "rms_in_w", "rms_ckv_w", "cos", "sin")
SEED, SEQ, GAIN = 4096, 4096, 2.4


def inputs_digest(d):
    h = hashlib.sha256()
    for k in KEYS:
        h.update(d[k].contiguous().numpy().tobytes())
    return h.hexdigest()


def main(out_path):
    so = glob.glob("oracle/_ref/_clusterfusion_ref*.so")[0]
    spec = importlib.util.spec_from_file_location("_clusterfusion_ref", so)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    d = D.make_inputs(SEQ, seed=SEED, out_gain=GAIN)
    c = [d[k].cuda() for k in KEYS]
    outs = []
    for _ in range(2):
        o = ref.deepseek_decoder_layer(*c)
        torch.cuda.synchronize()
        outs.append(o.cpu().reshape(-1))
    want, _, _ = D.deepseek_layer(**d)
    err = float((outs[0].float() - want.float().reshape(-1)).abs().max())
    rep = float((outs[0].float() - outs[1].float()).abs().max())
    print(f"reference kernel: |out| max {float(outs[0].float().abs().max()):.3f}, max|ref - oracle| {err:.4f}, max|launch 1 - launch 0| {rep:.4f}")
    np.savez(out_path, out=outs[0].numpy(), out_second_launch=outs[1].numpy(), inputs_sha256=inputs_digest(d), seed=SEED, seq_len=SEQ,
             out_gain=GAIN, max_abs_diff_vs_oracle=err)
    print("wrote", out_path)


if __name__ == "__main__":
    main(sys.argv[1])

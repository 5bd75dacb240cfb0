"""CPU check of the K tolerance the GPU suites use (tests/parity_helpers.py::close_k): it must accept what one or two fp16
rounding flips of a RoPE input can do to an output, and nothing more."""
import torch

from parity_helpers import close, close_k


def _rows(seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(64, 128, generator=g) * 2.5).half()


def test_close_k_accepts_input_rounding_flips_only():
    want = _rows()
    assert close_k(want.clone(), want)
    r = torch.hypot(want.float()[:, :64], want.float()[:, 64:]).repeat(1, 2)
    ulp = torch.exp2(torch.floor(torch.log2(r)) - 10)
    got = want.float().clone()
    got[3, 5] += 1.0 * ulp[3, 5]                       # one flip of one input
    got[7, 70] -= 2.0 * ulp[7, 70]                     # both inputs of a pair flipped the same way
    assert close_k(got, want)
    bad = want.float().clone()
    bad[3, 5] += 3.0 * ulp[3, 5]
    assert not close_k(bad, want)
    many = want.float() + 1.5 * ulp                    # every element more than 1 ulp off: far beyond 0.1 %
    assert not close_k(many, want)
    nan = want.float().clone(); nan[0, 0] = float("nan")
    assert not close_k(nan, want)


def test_close_k_small_outputs_of_large_pairs():
    """An output that came out small by cancellation still carries the absolute error of its large inputs: 2e-3 on a value of
    0.02 is 180 ulps OF THE OUTPUT but half an ulp of the pair magnitude 4.5 -- accepted; the plain 1e-3 check would not."""
    want = torch.zeros(1, 128, dtype=torch.float16)
    want[0, 0] = 0.02; want[0, 64] = 4.5
    got = want.float().clone(); got[0, 0] += 2e-3
    assert close_k(got, want) and not close(got, want)


def test_gptj_pairing():
    want = torch.zeros(1, 128, dtype=torch.float16)
    want[0, 0] = 0.02; want[0, 1] = 4.5                # GPT-J pairs (2i, 2i+1)
    got = want.float().clone(); got[0, 0] += 2e-3
    assert close_k(got, want, pairing="gptj") and not close_k(got, want, pairing="neox")

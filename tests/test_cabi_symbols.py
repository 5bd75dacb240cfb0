"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the pybind module exposes the reference's operator names, and misuse fails loudly.  No compute."""
import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    h = (ROOT / "include" / "clusterfusion_b200.h").read_text()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(cf_[a-z0-9_]+)\s*\(", h)))


def test_header_symbols_all_exported():
    from clusterfusion_b200 import cabi
    lib = cabi.load()
    syms = declared_symbols()
    assert set(syms) == set(cabi.EXPORTED_SYMBOLS), syms
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.cf_abi_version() == 3


def test_workspace_and_bytes_model():
    from clusterfusion_b200 import cabi
    assert cabi.workspace_bytes(4096, 1) >= 4096 * 4 + 5 * 4
    assert cabi.workspace_bytes(4096, 3) >= 3 * (4096 * 4 + 5 * 4)
    # batch >= 2 adds the exchange words of the grouped-query batched kernel, one set per chunk of 8 requests
    # (llama_decoder_gqa_batch_kernel.cuh: per-CTA QKV partials [160][768][8], states [160][8][528], per-group O partials [16][hidden][8])
    per_chunk = 8 * (160 * 768 * 8 + 160 * 8 * 528 + 16 * 4096 * 8)
    b1, b2, b8, b9 = (cabi.workspace_bytes(4096, b) for b in (1, 2, 8, 9))
    assert b2 - b1 >= per_chunk and b9 - b8 >= per_chunk and b8 - b2 < per_chunk
    assert cabi.workspace_bytes(8192, 8) > b8
    a = cabi.CfLlamaArgs(variant=cabi.CF_VARIANT_CHAT, hidden=4096, n_q_heads=32, n_kv_heads=32, head_dim=128, batch=1)
    # Llama-2-7B, kv 1K -> 151 036 928 B; kv 16K -> 402 695 168 B
    assert cabi.algorithmic_bytes(a, 1024) == 151_036_928 - 41_984 + (2 * 4096 * 3 + 2 * 2 * 4096 + 2 * 4 * 128)
    assert abs(cabi.algorithmic_bytes(a, 16384) - 402_695_168) < 64 * 1024


def test_argument_errors_without_gpu():
    from clusterfusion_b200 import cabi
    with pytest.raises(cabi.CfError) as e:
        cabi.launch(cabi.CfLlamaArgs(variant=9))
    assert e.value.code == -2 and "variant" in str(e.value)
    with pytest.raises(cabi.CfError) as e:
        cabi.launch(cabi.CfLlamaArgs(variant=0, head_dim=64))
    assert e.value.code == -3
    with pytest.raises(cabi.CfError) as e:
        cabi.launch(cabi.CfLlamaArgs(variant=0, head_dim=128, hidden=4096, n_q_heads=32, n_kv_heads=32, batch=1))
    assert e.value.code == -1      # NULL tensors


def test_pybind_surface_matches_reference_names():
    import clusterfusion
    for name in ("llama_decoder_layer", "llama_decoder_layer_sglang", "llama_decoder_layer_batch_decode_sglang", "rmsnorm"):
        assert callable(getattr(clusterfusion, name))
    doc = clusterfusion.llama_decoder_layer.__doc__
    assert doc.count("llama_decoder_layer(") >= 2        # 8-arg form + the README's 15-arg form
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        clusterfusion.llama_decoder_layer(*[torch.zeros(1) for _ in range(8)])
    with pytest.raises(TypeError):
        clusterfusion.llama_decoder_layer(torch.zeros(1))


def test_product_never_imports_oracle():
    """The product path must not route through the oracle (or any CPU fallback)."""
    for p in list((ROOT / "clusterfusion_b200").rglob("*.py")) + list((ROOT / "clusterfusion").rglob("*.py")):
        src = p.read_text()
        assert "oracle" not in re.sub(r'""".*?"""', "", src, flags=re.S), p


def test_new_argument_errors_without_gpu():
    """ABI v2 additions fail loudly on misuse (validated before anything touches a device)."""
    from clusterfusion_b200 import cabi
    base = dict(variant=1, head_dim=128, hidden=4096, n_q_heads=32, n_kv_heads=8, batch=1, x=16, residual_in=16, residual_out=32,
                w_qkv=16, w_o=16, rms_w=16, out=16, k_new=16, v_new=16, cos=16, sin=16, workspace=16)
    # rmsnorm: shape checks
    lib = cabi.load()
    assert lib.cf_rmsnorm_launch(16, 16, 16, 4, 100, 1e-6, 0, None) == -3          # hidden not a multiple of 16
    assert lib.cf_rmsnorm_launch(None, 16, 16, 4, 4096, 1e-6, 0, None) == -1
    assert lib.cf_tp_exchange_bytes(8192, 8) == 2 * 8 * 8192 * 8
    assert lib.cf_tp_exchange_bytes(0, 8) == 0
    # group counts beyond the group kernel's limit
    with pytest.raises(cabi.CfError) as e:
        cabi.launch(cabi.CfLlamaArgs(**{**base, "n_q_heads": 96, "n_kv_heads": 24, "hidden": 8192}))
    assert e.value.code in (-3, -5)


def test_ctypes_struct_mirrors_match_the_compiled_structs():
    import ctypes as C
    from clusterfusion_b200 import cabi
    lib = cabi.load()
    assert lib.cf_sizeof_llama_args() == C.sizeof(cabi.CfLlamaArgs)
    assert lib.cf_sizeof_ffn_args() == C.sizeof(cabi.CfFfnArgs)
    # spot-check field offsets against the header's declaration order (pointers are 8-byte aligned after 10 x 4 bytes)
    assert cabi.CfLlamaArgs.x.offset == 40 and cabi.CfLlamaArgs.workspace.offset == 40 + 18 * 8
    assert cabi.CfLlamaArgs.tp_peer.offset == cabi.CfLlamaArgs.workspace.offset + 8 + 3 * 4 + 4


def test_deepseek_boundary_without_gpu():
    """DeepSeek-MLA entry points: struct mirror, workspace size, shape / NULL validation before any device work, and the
    reference's operator name on the pybind module."""
    import ctypes as C
    import clusterfusion
    from clusterfusion_b200 import cabi
    lib = cabi.load()
    assert lib.cf_sizeof_deepseek_args() == C.sizeof(cabi.CfDeepseekArgs)
    assert cabi.CfDeepseekArgs.x.offset == 24 and cabi.CfDeepseekArgs.workspace.offset == 24 + 16 * 8
    # 129 flash-decode states of [16][512] floats dominate the workspace
    assert lib.cf_deepseek_workspace_bytes() >= 129 * 16 * 512 * 4 + 129 * 16 * 8 + 16 * 576 * 2 + (576 + 2048 + 8) * 4
    ptrs = {k: 16 for k in ("x", "w_q_nope", "w_q_pe", "w_uk", "w_kv_nope", "w_k_pe", "w_uv", "w_o", "ckv_cache", "rms_input_w",
                            "rms_ckv_w", "cos", "sin", "out", "workspace")}
    for bad, code in ((dict(hidden=4096), -3), (dict(n_heads=32), -3), (dict(seq_len=0), -3), (dict(x=None), -1),
                      (dict(ckv_cache=None), -1), (dict(out=24), -4)):
        with pytest.raises(cabi.CfError) as e:
            cabi.launch_deepseek(cabi.CfDeepseekArgs(**{**dict(hidden=2048, n_heads=16, seq_len=4096, eps=1e-6), **ptrs, **bad}))
        assert e.value.code == code, (bad, str(e.value))
    assert callable(clusterfusion.deepseek_decoder_layer)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        clusterfusion.deepseek_decoder_layer(*[torch.zeros(1) for _ in range(13)])

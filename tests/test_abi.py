"""The C-ABI library loads and exports every symbol include/gap_b200.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "gap_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gap_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    for required in ("gap_conv_gemm", "gap_conv_wgrad", "gap_bn_finalize", "gap_bn_act", "gap_bn_bwd_reduce",
                     "gap_bn_bwd_apply", "gap_bce_logits_const", "gap_gen_out_bwd", "gap_adam_flat",
                     "gap_pack_weights", "gap_last_error_string"):
        assert required in syms


def test_library_exports_every_declared_symbol(lib_path):
    handle = ctypes.CDLL(str(lib_path))
    missing = [s for s in _declared_symbols() if not hasattr(handle, s)]
    assert not missing, f"symbols declared in gap_b200.h but not exported: {missing}"


def test_python_binding_table_matches_header(lib_path):
    from gan_aug_pfa_b200 import _lib
    declared = set(_declared_symbols())
    bound = set(_lib._SIGNATURES)
    assert declared == bound, f"header-only: {declared - bound}; binding-only: {bound - declared}"
    h = _lib.lib()
    assert h.gap_version() >= 100
    assert isinstance(h.gap_last_error_string(), bytes)


def test_struct_sizes_are_plain_c():
    from gan_aug_pfa_b200 import _lib
    # pointers + ints only: sizes are stable multiples of 8 on LP64
    assert ctypes.sizeof(_lib.ConvGemmArgs) % 8 == 0
    assert ctypes.sizeof(_lib.WgradArgs) % 8 == 0


def test_bad_arguments_fail_loudly_without_a_gpu(lib_path):
    """Argument validation happens before any CUDA call, so it is testable on a CPU-only box."""
    from gan_aug_pfa_b200 import _lib
    h = _lib.lib()
    a = _lib.ConvGemmArgs()
    rc = h.gap_conv_gemm(ctypes.byref(a), None)
    assert rc == -1
    assert b"null" in h.gap_last_error_string()
    w = _lib.WgradArgs()
    assert h.gap_conv_wgrad(ctypes.byref(w), None) == -1


def test_product_path_has_no_cpu_fallback():
    """ops refuses CPU tensors instead of silently computing elsewhere."""
    import pytest
    import torch
    from gan_aug_pfa_b200 import ops
    x = torch.zeros(1, 4, 4, 64, dtype=torch.bfloat16)
    with pytest.raises(ValueError):
        ops._nhwc_view(x)
    src = (ROOT / "gan-aug-pfa_b200" / "ops.py").read_text() + (ROOT / "gan-aug-pfa_b200" / "pix2pix.py").read_text()
    assert "oracle" not in src, "product code must not import the oracle"
    assert "F.conv2d" not in src and "torch.nn.functional" not in src

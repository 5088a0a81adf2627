"""CPU: the C-ABI library loads without a GPU driver and exports every symbol include/bnerv_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "bnerv_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bnerv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported():
    from bnerv_b200 import _capi
    names = _declared()
    assert len(names) >= 14 and set(names) == set(_capi.EXPORTS)
    for n in names:
        assert hasattr(_capi.lib, n), n


def test_sizes_and_argument_errors_without_a_gpu():
    from bnerv_b200 import _capi
    lib = _capi.lib
    assert lib.bnerv_abi_version() == 4
    assert lib.bnerv_c8_numel(2, 135, 4, 5) == 2 * 144 * 20
    assert lib.bnerv_packed_weight_numel(135, 162, 3, 2) == 9 * 176 * 4 * 144
    assert lib.bnerv_packed_bias_numel(112, 2) == 448
    assert lib.bnerv_c8_numel(0, 1, 1, 1) == 0
    # argument validation happens before any CUDA call
    rc = lib.bnerv_conv_fused(None, 1, 1, 1, 1, None, None, 1, 3, 1, 0, None, None, None, None, None, None, None)
    assert rc == -1 and b"null" in lib.bnerv_last_error()
    one = ctypes.c_void_p(16)
    rc = lib.bnerv_conv_fused(one, 1, 8, 4, 4, one, one, 8, 5, 1, 0, None, None, None, one, None, None, None)
    assert rc == -2 and b"kernel size 5" in lib.bnerv_last_error()
    rc = lib.bnerv_pixel_shuffle(one, 1, 1, 1, 1, 0, one, None)
    assert rc == -1
    # backward entry points validate the same way
    assert lib.bnerv_wgrad_acc_numel(448, 135, 3) == 9 * 448 * 144
    assert lib.bnerv_conv_wgrad(None, one, 1, 8, 4, 4, 16, 3, one, None) == -1
    assert lib.bnerv_conv_wgrad(one, one, 1, 8, 4, 4, 24, 3, one, None) == -1 and b"multiple of 16" in lib.bnerv_last_error()
    assert lib.bnerv_conv_wgrad(one, one, 1, 8, 4, 4, 16, 5, one, None) == -2
    assert lib.bnerv_pack_conv_weight_dgrad(one, 8, 8, 2, 1, one, None) == -2
    assert lib.bnerv_head_bwd(one, one, 1, 3, 4, 4, None, one, one, None) == -1
    assert lib.bnerv_channel_sum(one, 1, 12, 4, 4, 0, one, None) == -1
    rc = lib.bnerv_conv_fused_ex(one, 1, 8, 4, 4, one, one, 8, 3, 1, 0, one, None, None, one, None, None, one, None)
    assert rc == -1 and b"out_deriv" in lib.bnerv_last_error()
    assert lib.bnerv_launch_count() == 0

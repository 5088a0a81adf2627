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
    # round-2 entry points: streaming kernels, split conv, paired linear layers, multi-tensor PTQ
    assert lib.bnerv_conv_stream(None, 1, 43, 8, 8, one, one, 43, 0, None, None, None, one, None, None) == -1
    assert lib.bnerv_conv_stream(one, 1, 43, 8, 8, one, one, 21, 0, None, None, None, one, None, None) == -2 and b"equal padded widths" in lib.bnerv_last_error()
    assert lib.bnerv_conv_stream(one, 1, 43, 8, 8, one, one, 43, 0, None, one, None, one, None, None) == -1       # g1p without beta / out_aff
    assert lib.bnerv_upconv_stream(one, 1, 12, 8, 8, one, one, 12, 1, one, one, one, one, None) == -2
    assert lib.bnerv_resblock_stream(one, one, 1, 64, 8, 8, one, one, one, one, 2, one, one, one, None) == -2
    assert lib.bnerv_resblock_stream_head(one, one, 1, 21, 8, 8, one, one, one, one, 2, one, one, one, one, 5, 4, one, None) == -2
    assert lib.bnerv_nerv_block_stream(one, 1, 12, 8, 8, one, one, 3, 3, 1, one, one, one, one, 12, 2, one, one, one, one, one, None) == -2
    assert lib.bnerv_conv_fused_split(one, 1, 48, 4, 4, one, one, 16, 3, 1, 0, None, None, None, one, None, None, 2, None) == -1
    assert lib.bnerv_linear_pair(None, 1, None) == -1
    assert lib.bnerv_ptq_quant_tensors(None, 0, 8, None, 0, None) == -1
    assert lib.bnerv_ptq_quant_tensors_scratch_bytes(None, 0) == 0
    assert lib.bnerv_launch_count() == 0

"""ctypes binding of libbnerv_b200.so (the C-ABI declared in include/bnerv_b200.h).

There is deliberately no fallback here: if the shared library is missing the import of this
module raises, and every wrapper raises BnervError on a non-zero return code.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libbnerv_b200.so")

E_BADARG, E_UNSUPPORTED, E_NODRIVER = -1, -2, -3
ACT_NONE, ACT_SIN, ACT_GELU, ACT_RELU, ACT_TANH01 = 0, 1, 2, 3, 4
ACT_CODES = {"none": ACT_NONE, "sin": ACT_SIN, "gelu": ACT_GELU, "relu": ACT_RELU, "tanh01": ACT_TANH01}

EXPORTS = [
    "bnerv_abi_version", "bnerv_last_error", "bnerv_launch_count", "bnerv_pack_conv_weight", "bnerv_conv_fused",
    "bnerv_conv_fused_f32", "bnerv_sft_affine", "bnerv_linear_act", "bnerv_nchw_to_c8", "bnerv_c8_to_nchw",
    "bnerv_pixel_shuffle", "bnerv_c8_numel", "bnerv_packed_weight_numel", "bnerv_packed_bias_numel",
    # backward of the cascade (ABI version 2)
    "bnerv_conv_fused_ex", "bnerv_conv_fused_split", "bnerv_pe_linear_pair", "bnerv_linear_pair", "bnerv_resblock_stream_head", "bnerv_upconv_stream", "bnerv_conv_stream", "bnerv_nerv_block_stream_head", "bnerv_head_bwd", "bnerv_pack_conv_weight_dgrad", "bnerv_conv_wgrad", "bnerv_wgrad_acc_numel",
    "bnerv_wgrad_finalize", "bnerv_bias_finalize", "bnerv_channel_sum", "bnerv_resblock_mid_bwd", "bnerv_block_front_bwd",
    "bnerv_unshuffle_c8", "bnerv_pack_conv_weight_q", "bnerv_frame_metrics", "bnerv_frame_metrics_scratch_doubles",
    "bnerv_pack_head_weight", "bnerv_head_conv3", "bnerv_nerv_block_fwd", "bnerv_head_conv1",
    "bnerv_ssim_stats", "bnerv_ssim_grad", "bnerv_ssim_scratch_floats",
    # post-training quantisation + Huffman statistics (ABI version 3)
    "bnerv_ptq_plan_tensor", "bnerv_ptq_quant_tensor", "bnerv_ptq_quant_tensors", "bnerv_ptq_quant_tensors_scratch_bytes", "bnerv_ptq_dequant_tensor", "bnerv_histogram_u8",
    "bnerv_huffman_code_lengths",
    # ConvNeXt encoder forward
    "bnerv_convnext_stage_fwd", "bnerv_convnext_stage_work_floats", "bnerv_nhwc_to_nchw",
    # one kernel per NeRVBlock for the narrow stages (ABI version 4)
    "bnerv_nerv_block_fused", "bnerv_resblock_fused", "bnerv_debug_set_buffer", "bnerv_nerv_block_stream",
    "bnerv_resblock_stream", "bnerv_bwd_set_status",
]


class BnervError(RuntimeError):
    def __init__(self, fn, code, text):
        super().__init__(f"{fn} failed with code {code}: {text}")
        self.code = code


class SftLayer(ctypes.Structure):
    """struct bnerv_sft_layer"""
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("ws0", "bs0", "ws1", "bs1", "wh0", "bh0", "wh1", "bh1", "g1p", "beta")] + \
               [("C", ctypes.c_int), ("Cp", ctypes.c_int)]


PTQ_MAX_CAND = 5


class PtqPlan(ctypes.Structure):
    """struct bnerv_ptq_plan"""
    _fields_ = [("n_cand", ctypes.c_int32), ("axis", ctypes.c_int32 * PTQ_MAX_CAND), ("groups", ctypes.c_int64 * PTQ_MAX_CAND),
                ("table_offset", ctypes.c_int64 * PTQ_MAX_CAND), ("table_floats", ctypes.c_int64),
                ("scratch_doubles", ctypes.c_int64)]


class LinearProblem(ctypes.Structure):
    """struct bnerv_linear_problem"""
    _fields_ = [("x", ctypes.c_void_p), ("w", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("y", ctypes.c_void_p),
                ("y_c8", ctypes.c_void_p), ("Cin", ctypes.c_int32), ("Cout", ctypes.c_int32), ("act", ctypes.c_int32),
                ("hw", ctypes.c_int32)]


class PtqJob(ctypes.Structure):
    """struct bnerv_ptq_job"""
    _fields_ = [("t", ctypes.c_void_p), ("shape", ctypes.c_int64 * (PTQ_MAX_CAND - 1)), ("ndim", ctypes.c_int32),
                ("reserved", ctypes.c_int32), ("quant", ctypes.c_void_p), ("new_t", ctypes.c_void_p), ("tables", ctypes.c_void_p),
                ("tables_f16", ctypes.c_void_p), ("err", ctypes.c_void_p), ("best", ctypes.c_void_p)]


class ConvNextBlock(ctypes.Structure):
    """struct bnerv_convnext_block"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("dw_w", "dw_b", "ln_w", "ln_b", "pw1_w", "pw1_b", "pw2_w", "pw2_b", "gamma")]


class ConvNextStage(ctypes.Structure):
    """struct bnerv_convnext_stage"""
    _fields_ = [(n, ctypes.c_void_p) for n in ("ln_in_w", "ln_in_b", "down_w", "down_b", "ln_out_w", "ln_out_b")] + \
               [("blocks", ctypes.POINTER(ConvNextBlock)), ("n_blocks", ctypes.c_int32), ("Cin", ctypes.c_int32),
                ("Cout", ctypes.c_int32), ("s", ctypes.c_int32)]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the bnerv_b200 CUDA extension is not built "
            "(run `python boosting-nerv_b200/build.py`); there is no fallback path.")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i, f = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    lib.bnerv_abi_version.restype = i
    lib.bnerv_last_error.restype = ctypes.c_char_p
    lib.bnerv_launch_count.restype = ctypes.c_uint64
    lib.bnerv_pack_conv_weight.argtypes = [vp, vp, i, i, i, i, vp, vp, vp]
    lib.bnerv_conv_fused.argtypes = [vp, i, i, i, i, vp, vp, i, i, i, i, vp, vp, vp, vp, vp, vp, vp]
    lib.bnerv_pack_conv_weight_q.argtypes = [vp, vp, i, vp, vp, i, i, i, i, i, i, vp, vp, vp]
    lib.bnerv_frame_metrics.argtypes = [vp, vp, i, ctypes.c_size_t, vp, vp, vp]
    lib.bnerv_frame_metrics_scratch_doubles.argtypes = [i]
    lib.bnerv_frame_metrics_scratch_doubles.restype = ctypes.c_size_t
    lib.bnerv_pack_head_weight.argtypes = [vp, i, i, vp, vp]
    lib.bnerv_head_conv3.argtypes = [vp, i, i, i, i, vp, vp, i, i, vp, vp]
    lib.bnerv_nerv_block_fwd.argtypes = [vp, i, i, i, i, vp, vp, i, i, i, vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.bnerv_nerv_block_fused.argtypes = [vp, i, i, i, i, vp, vp, i, i, i, vp, vp, vp, vp, i, i, vp, vp, vp, vp, vp, vp]
    lib.bnerv_resblock_fused.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp, i, vp, vp, vp, vp]
    lib.bnerv_debug_set_buffer.argtypes = [vp, i]
    lib.bnerv_bwd_set_status.argtypes = [vp]
    lib.bnerv_nerv_block_stream.argtypes = lib.bnerv_nerv_block_fused.argtypes
    lib.bnerv_resblock_stream.argtypes = lib.bnerv_resblock_fused.argtypes
    lib.bnerv_head_conv1.argtypes = [vp, i, i, i, i, vp, vp, i, i, vp, vp]
    lib.bnerv_ssim_stats.argtypes = [vp, vp, i, i, i, f, f, vp, vp]
    lib.bnerv_ssim_grad.argtypes = [vp, vp, i, i, i, f, f, vp, vp, i, vp, vp]
    lib.bnerv_ssim_scratch_floats.argtypes = [i, i, i]
    lib.bnerv_ssim_scratch_floats.restype = ctypes.c_size_t
    lib.bnerv_convnext_stage_fwd.argtypes = [ctypes.POINTER(ConvNextStage), vp, i, i, i, i, vp, vp, vp]
    lib.bnerv_convnext_stage_work_floats.argtypes = [i, i, i, i, i]
    lib.bnerv_convnext_stage_work_floats.restype = ctypes.c_size_t
    lib.bnerv_nhwc_to_nchw.argtypes = [vp, i, i, i, i, vp, vp]
    lib.bnerv_ptq_plan_tensor.argtypes = [vp, i, ctypes.POINTER(PtqPlan)]
    lib.bnerv_ptq_quant_tensor.argtypes = [vp, vp, i, i, vp, vp, vp, vp, vp, vp, vp]
    lib.bnerv_resblock_stream_head.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp, i, vp, vp, vp, vp, i, i, vp, vp]
    lib.bnerv_nerv_block_stream_head.argtypes = [vp, i, i, i, i, vp, vp, vp, vp, vp, vp, i, vp, vp, vp, vp, vp, vp, i, i, vp, vp]
    lib.bnerv_conv_stream.argtypes = [vp, i, i, i, i, vp, vp, i, i, vp, vp, vp, vp, vp, vp]
    lib.bnerv_upconv_stream.argtypes = [vp, i, i, i, i, vp, vp, i, i, vp, vp, vp, vp, vp]
    lib.bnerv_pe_linear_pair.argtypes = [vp, i, vp, i, vp, vp]
    lib.bnerv_linear_pair.argtypes = [vp, i, vp]
    lib.bnerv_ptq_quant_tensors.argtypes = [vp, i, i, vp, ctypes.c_size_t, vp]
    lib.bnerv_ptq_quant_tensors_scratch_bytes.argtypes = [vp, i]
    lib.bnerv_ptq_quant_tensors_scratch_bytes.restype = ctypes.c_size_t
    lib.bnerv_ptq_dequant_tensor.argtypes = [vp, vp, i, i, vp, vp, i, vp, vp]
    lib.bnerv_histogram_u8.argtypes = [vp, ctypes.c_size_t, vp, vp]
    lib.bnerv_huffman_code_lengths.argtypes = [vp, i, vp]
    lib.bnerv_conv_fused_ex.argtypes = [vp, i, i, i, i, vp, vp, i, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.bnerv_conv_fused_split.argtypes = [vp, i, i, i, i, vp, vp, i, i, i, i, vp, vp, vp, vp, vp, vp, i, vp]
    lib.bnerv_head_bwd.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, vp]
    lib.bnerv_pack_conv_weight_dgrad.argtypes = [vp, i, i, i, i, vp, vp]
    lib.bnerv_conv_wgrad.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp]
    lib.bnerv_wgrad_acc_numel.argtypes = [i, i, i]
    lib.bnerv_wgrad_acc_numel.restype = ctypes.c_size_t
    lib.bnerv_wgrad_finalize.argtypes = [vp, i, i, i, i, vp, i, vp, vp]
    lib.bnerv_bias_finalize.argtypes = [vp, i, i, vp, i, vp, vp]
    lib.bnerv_channel_sum.argtypes = [vp, i, i, i, i, i, vp, vp]
    lib.bnerv_resblock_mid_bwd.argtypes = [vp, vp, vp, vp, i, i, i, i, vp, vp, vp, vp, vp]
    lib.bnerv_block_front_bwd.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, vp, vp, vp, vp, vp, vp]
    lib.bnerv_unshuffle_c8.argtypes = [vp, i, i, i, i, i, vp, vp, vp]
    lib.bnerv_conv_fused_f32.argtypes = [vp, i, i, i, i, vp, vp, i, i, i, i, vp, vp, vp, i, vp, vp, vp]
    lib.bnerv_sft_affine.argtypes = [vp, i, vp, i, i, vp]
    lib.bnerv_linear_act.argtypes = [vp, i, i, vp, vp, i, i, vp, vp]
    lib.bnerv_nchw_to_c8.argtypes = [vp, i, i, i, i, vp, vp]
    lib.bnerv_c8_to_nchw.argtypes = [vp, i, i, i, i, vp, vp]
    lib.bnerv_pixel_shuffle.argtypes = [vp, i, i, i, i, i, vp, vp]
    for name in ("bnerv_c8_numel",):
        getattr(lib, name).argtypes = [i, i, i, i]
        getattr(lib, name).restype = ctypes.c_size_t
    lib.bnerv_packed_weight_numel.argtypes = [i, i, i, i]
    lib.bnerv_packed_weight_numel.restype = ctypes.c_size_t
    lib.bnerv_packed_bias_numel.argtypes = [i, i]
    lib.bnerv_packed_bias_numel.restype = ctypes.c_size_t
    missing = [name for name in EXPORTS if not hasattr(lib, name)]
    if missing:
        raise ImportError(f"{LIB_PATH} is stale: missing {missing} (re-run `python boosting-nerv_b200/build.py`)")
    return lib


lib = _load()


def check(fn_name, rc):
    if rc != 0:
        raise BnervError(fn_name, rc, lib.bnerv_last_error().decode("utf-8", "replace"))


def launch_count():
    return int(lib.bnerv_launch_count())


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else ctypes.c_void_p(t.data_ptr())

"""bnerv_b200 — B200-native (sm_100a) Boosting-NeRV conditional-decoder hot path.

Importing this package does not load the CUDA extension; `bnerv_b200.ops` / `bnerv_b200.engine` do,
and raise if libbnerv_b200.so has not been built (no fallback).
"""
from .config import make_args, preset, solve_fc_dim, tiny_args  # noqa: F401
from .models import ENeRV_Boost, HNeRV_Boost, NeRV_Boost  # noqa: F401
from .stream import decode_to_host, evaluate_metrics, evaluate_psnr  # noqa: F401

__all__ = ["NeRV_Boost", "ENeRV_Boost", "HNeRV_Boost", "decode_to_host", "evaluate_metrics", "evaluate_psnr", "make_args", "preset", "solve_fc_dim", "tiny_args"]

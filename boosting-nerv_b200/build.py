"""Build libbnerv_b200.so in-tree with nvcc for sm_100a (no torch headers, no libcuda link).

Usage: python boosting-nerv_b200/build.py [--force]
The library links cudart statically and resolves cuTensorMapEncodeTiled through
cudaGetDriverEntryPoint at run time, so it dlopens on a box without a GPU driver.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbnerv_b200.so")
SOURCES = ["capi.cu", "ops.cu", "conv_tc.cu", "block_fused.cu", "block_stream.cu", "block_stream32.cu", "conv_stream.cu", "conv_wgrad.cu", "bwd_ops.cu", "loss_ops.cu", "ptq_ops.cu", "encoder_ops.cu"]
FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3",
         "-std=c++17", "--cudart", "static"]


def _source_hash():
    """Content hash of everything the library is built from (mtimes do not survive a snapshot copy to another box)."""
    import hashlib
    h = hashlib.sha256(" ".join(FLAGS + SOURCES).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "bnerv_b200.h")]
    for d in deps:
        with open(d, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB) or not os.path.exists(LIB + ".srchash"):
        return True
    with open(LIB + ".srchash") as fh:
        return fh.read().strip() != _source_hash()


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    with open(LIB + ".srchash", "w") as fh:
        fh.write(_source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

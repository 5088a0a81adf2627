"""Pipelined host-to-host decode: the throughput form of the reference's evaluate() loop.

The reference decodes frame by frame and pulls every result to the host synchronously (`model(...)` with its
`torch.cuda.synchronize()` at model_nerv.py:58-59, then `.cpu()` in the metric code, train_nerv_all.py:482-505), so the
25 MB device-to-host copy of a 1080p f32 frame (~1 ms over PCIe) sits on the critical path of every frame.  Here the
same three steps run as a three-stage pipeline on two streams:

    compute stream : H2D of (embedding, norm_idx) -> one CUDA-graph replay -> 8 us device copy into a staging slot
    copy stream    : D2H of the staging slot into the caller's pinned buffer

with `depth` staging slots guarded by events, so frame i's read-back overlaps frame i+1's decode.  Results are
bit-identical to model.forward()/forward_decoder() frame by frame (same graph, same kernels).
"""
import weakref

import torch

_COPY_STREAMS = weakref.WeakKeyDictionary()


def decode_to_host(model, norm_idx_host, out_host, embed_host=None, batch=1, depth=3, ring=False):
    """Decode frames [0, N) into ``out_host`` ([N,3,H,W] f32, ideally pinned).  With ``ring=True`` ``out_host`` may hold
    fewer than N frames (a multiple of ``batch``) and is used cyclically - frame i lands in slot i % len(out_host) - for
    consumers that drain the buffer while decoding continues.

    norm_idx_host: [N] float64 host tensor ((i+1)/n_frames, hnerv_utils.py:47); embed_host: [N,16,h,w] f32 host tensor
    for HNeRV_Boost, None for NeRV_Boost / ENeRV_Boost.  Returns when every frame has landed in ``out_host``."""
    if norm_idx_host.is_cuda or out_host.is_cuda or (embed_host is not None and embed_host.is_cuda):
        raise ValueError("decode_to_host takes HOST tensors (pinned for asynchronous copies)")
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("bnerv_b200: the decode path runs only on a CUDA (sm_100a) device")
    n = norm_idx_host.shape[0]
    slots = out_host.shape[0]
    assert (slots == n or (ring and slots % batch == 0 and slots > 0)) and (embed_host is None or embed_host.shape[0] == n)
    is_h = embed_host is not None
    compute = torch.cuda.current_stream(dev)
    copy = _COPY_STREAMS.get(model)              # kept outside the module: deepcopy(model) must stay a plain parameter copy
    if copy is None or copy.device != dev:
        copy = _COPY_STREAMS[model] = torch.cuda.Stream(dev)
    stages, ready, done = {}, {}, {}
    model.engine().sync_weights()                # model.decode() replays captured graphs: re-check the weights once per call
    with torch.no_grad():
        for j, lo in enumerate(range(0, n, batch)):
            sl = slice(lo, min(lo + batch, n))
            k = j % depth
            t = norm_idx_host[sl].to(dev, non_blocking=True)
            img = model.decode(embed_host[sl].to(dev, non_blocking=True), t) if is_h else model.decode(t)
            key = (k, tuple(img.shape))
            if key not in stages:
                stages[key] = torch.empty_like(img)
                ready[key], done[key] = torch.cuda.Event(), None
            if done[key] is not None:
                compute.wait_event(done[key])          # the slot's previous read-back (depth frames ago) has finished
            stages[key].copy_(img)
            ready[key].record(compute)
            with torch.cuda.stream(copy):
                copy.wait_event(ready[key])
                lo_s = lo % slots
                out_host[lo_s:lo_s + (sl.stop - sl.start)].copy_(stages[key], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                done[key] = ev
    copy.synchronize()
    return out_host


def evaluate_psnr(model, norm_idx_host, gt_host, embed_host=None, batch=1):
    """Mean PSNR of this rank's frames against ground truth, then over all ranks - the metric half of the reference's
    evaluate() (train_nerv_all.py:482-505, 554-556) without a host synchronisation per frame: ground-truth frames are
    copied H2D ahead of the decode, mse/psnr come from `bnerv_frame_metrics` on the device (psnr_fn_single,
    hnerv_utils.py:400-403), the per-frame values are summed on the device and ONE all_reduce(SUM) of (sum, count)
    runs at the end (hnerv_utils.py:213-229) when torch.distributed is initialised.  Returns (mean_psnr, n_frames_total)."""
    return evaluate_metrics(model, norm_idx_host, gt_host, embed_host, batch, with_msssim=False)[:2]


def evaluate_metrics(model, norm_idx_host, gt_host, embed_host=None, batch=1, with_msssim=True):
    """evaluate_psnr plus the mean MS-SSIM (msssim_fn_single, hnerv_utils.py:406-408) from the device kernels of
    bnerv_b200.losses.  Returns (mean_psnr, n_frames_total, mean_msssim or None); one all_reduce per metric."""
    from . import ops
    from .shard import reduce_metric
    dev = next(model.parameters()).device
    if dev.type != "cuda":
        raise RuntimeError("bnerv_b200: the decode path runs only on a CUDA (sm_100a) device")
    n = norm_idx_host.shape[0]
    assert gt_host.shape[0] == n and gt_host.dtype == torch.float32
    is_h = embed_host is not None
    total = torch.zeros(2, dtype=torch.float64, device=dev)
    if with_msssim:
        from .losses import ms_ssim
    model.engine().sync_weights()                # see decode_to_host: weights changed since the last call invalidate the graphs
    with torch.no_grad():
        for lo in range(0, n, batch):
            sl = slice(lo, min(lo + batch, n))
            gt = gt_host[sl].to(dev, non_blocking=True)
            t = norm_idx_host[sl].to(dev, non_blocking=True)
            img = model.decode(embed_host[sl].to(dev, non_blocking=True), t) if is_h else model.decode(t)
            total[0] += ops.frame_metrics(img, gt)[:, 2].double().sum()
            if with_msssim:
                total[1] += ms_ssim(img, gt, data_range=1, size_average=False).double().sum()
    sums = total.tolist()
    psnr, cnt = reduce_metric(sums[0], n, device=dev)
    msssim = reduce_metric(sums[1], n, device=dev)[0] if with_msssim else None
    return psnr, cnt, msssim

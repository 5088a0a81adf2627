"""CPU restatement of pytorch_msssim==0.2.1 (the version pinned in the reference's requirements.txt) - TEST INFRASTRUCTURE.

The reference's loss (hnerv_utils.py:338-395: `ssim(pred, target, data_range=1, size_average=False)` and
`ms_ssim(...)` inside the 'Fusion*' losses) calls this third-party package, which is neither vendored in the reference
tree nor installed here.  PARITY UNPINNED: there are no golden vectors of the package to check this file against; it
restates the package's published algorithm (Gaussian 11-tap window sigma 1.5, separable 'valid' filtering, K = (0.01,
0.03), relu on the per-channel cs / ssim means inside ms_ssim only (ssim's nonnegative_ssim defaults to False), five levels with weights (0.0448, 0.2856, 0.3001, 0.2363, 0.1333),
2x2 average pooling with padding = size % 2 between levels, product of powers).  Only tests/ may import it.
"""
import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def gauss_1d(size=11, sigma=1.5, dtype=torch.float32):
    coords = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def gaussian_filter(x, win):
    c = x.shape[1]
    out = x
    for i, s in enumerate(x.shape[2:]):
        if s >= win.shape[-1]:
            w = win.view(1, 1, 1, -1).repeat(c, 1, 1, 1)
            out = F.conv2d(out, w.transpose(2 + i, -1), stride=1, padding=0, groups=c)
    return out


def ssim_stats(x, y, data_range=1.0, k=(0.01, 0.03)):
    """-> (ssim_per_channel [B,C], cs_per_channel [B,C]) of one level."""
    win = gauss_1d(dtype=x.dtype).to(x.device)
    c1, c2 = (k[0] * data_range) ** 2, (k[1] * data_range) ** 2
    mu1, mu2 = gaussian_filter(x, win), gaussian_filter(y, win)
    s11 = gaussian_filter(x * x, win) - mu1 * mu1
    s22 = gaussian_filter(y * y, win) - mu2 * mu2
    s12 = gaussian_filter(x * y, win) - mu1 * mu2
    cs_map = (2 * s12 + c2) / (s11 + s22 + c2)
    ssim_map = ((2 * mu1 * mu2 + c1) / (mu1 * mu1 + mu2 * mu2 + c1)) * cs_map
    return ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)


def ssim(x, y, data_range=1.0, size_average=False):
    # nonnegative_ssim=False is the package default in 0.2.1: no clamp on the single-scale path (ms_ssim clamps below)
    v = ssim_stats(x, y, data_range)[0]
    return v.mean() if size_average else v.mean(1)


def ms_ssim(x, y, data_range=1.0, size_average=False):
    assert min(x.shape[-2:]) > (11 - 1) * 2 ** 4, "image too small for 5 levels"
    w = torch.tensor(WEIGHTS, dtype=x.dtype, device=x.device)
    mcs = []
    for i in range(len(WEIGHTS)):
        s, cs = ssim_stats(x, y, data_range)
        if i < len(WEIGHTS) - 1:
            mcs.append(torch.relu(cs))
            pad = [sz % 2 for sz in x.shape[2:]]
            x = F.avg_pool2d(x, kernel_size=2, padding=pad)
            y = F.avg_pool2d(y, kernel_size=2, padding=pad)
    vals = torch.stack(mcs + [torch.relu(s)], dim=0)
    out = torch.prod(vals ** w.view(-1, 1, 1), dim=0)
    return out.mean() if size_average else out.mean(1)

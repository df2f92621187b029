"""The mapper's photometric / geometric / semantic loss, restated from the reference with stock
torch ops (include/loss_utils.h:27-131; call site src/gaussian_mapper.cpp:707-724).  Out of
scope for the CUDA work: both arms of every comparison use it
unchanged; it only produces dL/dpixel for the rasterizer backward."""
import math

import torch
import torch.nn.functional as F


def l1_loss(x, gt):
    return (x - gt).abs().mean()


def psnr(img1, img2):
    mse = ((img1 - img2) ** 2).mean()
    return 10.0 * torch.log10(1.0 / mse)


def cosine_similarity(lf, gt):
    """mean over pixels of cos(lf[:,p], gt[:,p]) (loss_utils.h:36-40)."""
    c = lf.shape[0]
    return F.cosine_similarity(lf.reshape(c, -1), gt.reshape(c, -1), dim=0).mean()


_windows = {}


def _window(channel, device, window_size=11, sigma=1.5):
    key = (channel, str(device), window_size)
    if key not in _windows:
        g = torch.tensor([math.exp(-((x - window_size // 2) ** 2) / (2.0 * sigma * sigma)) for x in range(window_size)],
                         dtype=torch.float32, device=device)
        g = (g / g.sum()).unsqueeze(1)
        _windows[key] = (g @ g.t()).unsqueeze(0).unsqueeze(0).expand(channel, 1, window_size, window_size).contiguous()
    return _windows[key]


def ssim(img1, img2, window_size=11):
    channel = img1.size(-3)
    w = _window(channel, img1.device, window_size)
    pad = window_size // 2
    if img1.dim() == 3:
        img1, img2 = img1.unsqueeze(0), img2.unsqueeze(0)
    mu1 = F.conv2d(img1, w, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, w, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(img1 * img1, w, padding=pad, groups=channel) - mu1_sq
    s2 = F.conv2d(img2 * img2, w, padding=pad, groups=channel) - mu2_sq
    s12 = F.conv2d(img1 * img2, w, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    return (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()


def mapping_loss(image, lf, depth, gt_image, gt_lf, gt_depth, lambda_dssim=0.2, faithful_sign=True):
    """0.8*L1 + 0.2*(1-SSIM) + cos_sim(lf) + L1(depth)  (src/gaussian_mapper.cpp:716-721).
    The reference ADDS the mean cosine similarity; `faithful_sign=False`
    uses 1 - cos instead, which is what the PSNR-parity tests need to converge."""
    Ll1 = l1_loss(image, gt_image)
    loss = (1.0 - lambda_dssim) * Ll1 + lambda_dssim * (1.0 - ssim(image, gt_image))
    sim = cosine_similarity(lf, gt_lf)
    loss = loss + (sim if faithful_sign else (1.0 - sim))
    return loss + l1_loss(depth, gt_depth)

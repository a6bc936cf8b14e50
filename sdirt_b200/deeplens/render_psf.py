"""Spatially varying dual-pixel rendering (deeplens/render_psf.py:76-188) on the CUDA engine.

One kernel serves the three reference entry points: they differ only in tensor layout."""
import torch

from .. import _engine as E


def _prep(input, psf, kernel_size):
    if input.dim() < 4:
        input = input.unsqueeze(0)
    img = input.float().contiguous()
    b, c, h, w = img.shape
    if psf.dtype not in (torch.float16, torch.float32):
        psf = psf.float()
    psf = psf.reshape(b, h, w, 2, kernel_size, kernel_size).contiguous()
    return input.dtype, img, psf


def local_psf_render_fast(input, psf, kernel_size=11, val=False):
    """(rl, rr) = per-pixel convolution of `input` [B,C,H,W] with its left / right PSFs `psf` [B,H,W,2,ks,ks];
    fp16 products and fp16-rounded sums exactly as the reference's half() path (render_psf.py:120-155)."""
    orig, img, psf = _prep(input, psf, kernel_size)
    rl, rr = E.render_local_psf(img, psf, kernel_size)
    return rl.to(orig), rr.to(orig)


def local_psf_render(input, psf, kernel_size=11, val=False):
    """(rl, rr) like local_psf_render_fast: the reference's older formulation of the same half() arithmetic (render_psf.py:76-118;
    its docstring still describes a single-PSF signature, its body reshapes `psf` to [-1, 2, ks, ks] and returns the pair)."""
    return local_psf_render_fast(input, psf, kernel_size, val)


def local_dp_psf_render(input, dp_psf, kernel_size=21):
    """[N, 2C, H, W] = cat(left, right) (render_psf.py:157-188).  The reference does NOT cast this variant to half: the
    arithmetic runs in the promoted dtype of `input` and `dp_psf` -- float32 unless both are half."""
    if input.dtype == torch.float16 and dp_psf.dtype == torch.float16:
        rl, rr = local_psf_render_fast(input, dp_psf, kernel_size)
        return torch.cat([rl, rr], dim=1)
    b, c, h, w = input.shape
    img = input.float().contiguous()
    psf = dp_psf.float().reshape(b, h, w, 2, kernel_size, kernel_size).contiguous()
    rl, rr = E.render_local_psf_f32(img, psf, kernel_size)
    return torch.cat([rl, rr], dim=1)

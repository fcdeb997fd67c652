"""Numerics study (CPU, not a test): can the two correction products of the split-precision GEMM run in fp8?

The product path computes every contraction as  A_hi·W_hi + A_lo·W_hi + A_hi·W_lo  with fp16 planes.
The corrections are 2^-11 of the result, so they tolerate a much coarser operand format.  This script emulates, inside the CPU
oracle (F.conv2d of the weighted 3x3/1x1 convs only; attention and the small linears stay exact),

    fp16      : A_hi·W_hi                                            (1 MMA unit)
    fp16x3    : A_hi·W_hi + A_lo·W_hi + A_hi·W_lo, fp16 planes       (3 units; today's default)
    fp16+f8   : A_hi·W_hi + e4m3(A_lo·2^13)·e4m3(W_hi·2^b1)·2^-S + e4m3(A_hi·2^2)·e4m3(W_lo·2^(S-2))·2^-S,  S = 13 + b1,
                b1 = floor(log2(448 / max|W|))                      (2 units: e4m3 MMAs run at twice the fp16 rate)

and prints the max-abs error of the denoiser output D against the fp32 oracle on the de-zeroed nets.

    python tests/study_fp8_corrections.py [net ...]
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import edm_oracle as O   # noqa: E402

_real_conv2d = F.conv2d
E4M3_MAX = 448.0


def split16(x):
    hi = x.to(torch.float16).to(torch.float32)
    lo = (x - hi).to(torch.float16).to(torch.float32)
    return hi, lo


def q8(x, log2_scale):
    s = 2.0 ** log2_scale
    return (x * s).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).to(torch.float32) / s


SH_A16, SH_LO8, SH_HI8 = 6, 13, 2          # csrc/ops.h DS_F8_SH_*


def f16_scaled(x, sh):
    """fp16 plane of x * 2^sh (saturating), returned unscaled."""
    return (x * 2.0 ** sh).clamp(-65504.0, 65504.0).to(torch.float16).to(torch.float32) / 2.0 ** sh


def make_conv(mode):
    def conv2d(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1):
        if groups != 1 or mode == 'fp32':
            return _real_conv2d(x, w, bias, stride, padding, dilation, groups)
        if mode == 'fp16+f8':
            # exactly the operand formats of ds_gemm_desc.f8 (csrc/ops.h) / gemm_desc.pack_conv_weight_f8
            b1 = int(torch.floor(torch.log2(E4M3_MAX / w.abs().max())).item())
            S_ = SH_LO8 + b1
            xh = f16_scaled(x, SH_A16)
            xl8 = q8(x - xh, SH_LO8)
            xh8 = q8(xh, SH_HI8)
            wh = f16_scaled(w, S_ - SH_A16)
            wh8 = q8(wh, b1)
            wl8 = q8(w - wh, S_ - SH_HI8)
            y = _real_conv2d(xl8, wh8, None, stride, padding) + _real_conv2d(xh8, wl8, None, stride, padding)
            y = y + _real_conv2d(xh, wh, None, stride, padding)
        else:
            xh, xl = split16(x)
            wh, wl = split16(w)
            y = _real_conv2d(xh, wh, None, stride, padding)
            if mode == 'fp16x3':
                y = y + _real_conv2d(xl, wh, None, stride, padding) + _real_conv2d(xh, wl, None, stride, padding)
            elif mode != 'fp16':
                raise ValueError(mode)
        if bias is not None:
            y = y + bias.reshape(1, -1, 1, 1)
        return y
    return conv2d


def run(name, batch=2, sigmas=(80.0, 2.0, 0.05)):
    P, S = O.make_net(name, seed=0, dezero=True)
    net = O.OracleNet(P, S)
    res = S['img_resolution']
    lat = O.stacked_randn(range(batch), (S['img_channels'], res, res))
    labels = None
    if S['label_dim']:
        labels = torch.eye(S['label_dim'])[torch.arange(batch) % S['label_dim']]
    out = {}
    for sg in sigmas:
        x = lat * sg
        ref = None
        for mode in ('fp32', 'fp16', 'fp16x3', 'fp16+f8'):
            O.F.conv2d = make_conv(mode)
            t0 = time.time()
            with torch.no_grad():
                D = net(x, torch.tensor(sg), class_labels=labels)
            if mode == 'fp32':
                ref = D
            else:
                out.setdefault(mode, []).append((D - ref).abs().max().item())
            print(f'  {name} sigma={sg:<6} {mode:8s} max|D|={D.abs().max():.3f} err={0.0 if ref is D else (D - ref).abs().max().item():.3e} ({time.time() - t0:.1f}s)',
                  flush=True)
    O.F.conv2d = _real_conv2d
    return out


if __name__ == '__main__':
    torch.set_num_threads(os.cpu_count())
    for n in (sys.argv[1:] or ['tiny_song', 'cifar10']):
        r = run(n)
        print(n, {k: f'{max(v):.2e}' for k, v in r.items()})

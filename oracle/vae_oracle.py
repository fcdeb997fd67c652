"""ORACLE — test infrastructure, not product code.

CPU restatement of the latent-diffusion first-stage DECODER, i.e. what the reference runs after sampling
(sample.py:299 `net.model.decode_first_stage(images)`):

    z = z / scale_factor                                  ddpm.py:714
    z = post_quant_conv(z)                                autoencoder.py AutoencoderKL.decode (1x1 conv, embed_dim -> z_channels)
    x = Decoder(z)                                        modules/diffusionmodules/model.py:462-569

written functionally over a flat parameter dict named like `first_stage_model.state_dict()` (`post_quant_conv.*`, `decoder.*`).
Pinned against the real reference classes by tests/golden/ref_vae.npz (oracle/gen_vae_golden.py).  Only tests/ may import this.
Paths are relative to /root/reference/diff-solvers-main/models/ldm/.
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

CONFIGS = {
    # models/ldm/configs/stable-diffusion/v1-inference.yaml:46-65 (first_stage_config.ddconfig); latents 64x64 -> images 512x512
    'sd_vae': dict(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, z_channels=4, embed_dim=4, scale_factor=0.18215),
    # reduced nets with the same structure (mid attention, channel-changing blocks with nin_shortcut, upsampling convs)
    'tiny_vae': dict(ch=64, out_ch=3, ch_mult=(1, 2), num_res_blocks=1, z_channels=4, embed_dim=4, scale_factor=0.18215),
    'wide_vae': dict(ch=64, out_ch=3, ch_mult=(1, 1, 1), num_res_blocks=0, z_channels=4, embed_dim=4, scale_factor=0.18215),
}


def structure(cfg):
    """Module list of Decoder.__init__ (model.py:462-531) in execution order:
    [('conv', name, cin, cout) | ('res', name, cin, cout) | ('attn', name, c) | ('up', name, c)], final channels."""
    ch, mult, nrb = cfg['ch'], cfg['ch_mult'], cfg['num_res_blocks']
    nres = len(mult)
    block_in = ch * mult[nres - 1]
    mods = [('conv', 'decoder.conv_in', cfg['z_channels'], block_in),
            ('res', 'decoder.mid.block_1', block_in, block_in), ('attn', 'decoder.mid.attn_1', block_in),
            ('res', 'decoder.mid.block_2', block_in, block_in)]
    for lvl in reversed(range(nres)):
        block_out = ch * mult[lvl]
        for i in range(nrb + 1):
            mods.append(('res', f'decoder.up.{lvl}.block.{i}', block_in, block_out))
            block_in = block_out
        if lvl != 0:
            mods.append(('up', f'decoder.up.{lvl}.upsample', block_in))
    return mods, block_in


def make_params(name, seed=0):
    """Random-init parameters in torch's default Conv2d / GroupNorm init, drawn in module order (values only matter for parity;
    the golden fixture pins the FORWARD against the reference classes loaded with these very tensors)."""
    cfg = dict(CONFIGS[name])
    g = torch.Generator().manual_seed(seed)
    P = OrderedDict()

    def conv(n, cin, cout, k):
        bound = 1.0 / np.sqrt(cin * k * k)
        P[n + '.weight'] = (torch.rand(cout, cin, k, k, generator=g) * 2 - 1) * bound
        P[n + '.bias'] = (torch.rand(cout, generator=g) * 2 - 1) * bound

    def norm(n, c):
        P[n + '.weight'] = 1.0 + 0.2 * torch.randn(c, generator=g)
        P[n + '.bias'] = 0.1 * torch.randn(c, generator=g)

    conv('post_quant_conv', cfg['embed_dim'], cfg['z_channels'], 1)
    mods, c_end = structure(cfg)
    for m in mods:
        if m[0] == 'conv':
            conv(m[1], m[2], m[3], 3)
        elif m[0] == 'res':
            _, n, cin, cout = m
            norm(n + '.norm1', cin)
            conv(n + '.conv1', cin, cout, 3)
            norm(n + '.norm2', cout)
            conv(n + '.conv2', cout, cout, 3)
            if cin != cout:
                conv(n + '.nin_shortcut', cin, cout, 1)
        elif m[0] == 'attn':
            _, n, c = m
            norm(n + '.norm', c)
            for k in ('q', 'k', 'v', 'proj_out'):
                conv(n + '.' + k, c, c, 1)
        elif m[0] == 'up':
            conv(m[1] + '.conv', m[2], m[2], 3)
    norm('decoder.norm_out', c_end)
    conv('decoder.conv_out', c_end, cfg['out_ch'], 3)
    return P, cfg


def _gn(P, n, x):
    """Normalize = GroupNorm(32 groups, eps=1e-6, affine)  (model.py:37-38)."""
    return F.group_norm(x, 32, P[n + '.weight'].to(x.dtype), P[n + '.bias'].to(x.dtype), 1e-6)


def _conv(P, n, x, pad):
    return F.conv2d(x, P[n + '.weight'].to(x.dtype), P[n + '.bias'].to(x.dtype), padding=pad)


def _swish(x):
    return x * torch.sigmoid(x)                                     # model.py:32-34 nonlinearity


def _res(P, n, x, cin, cout):
    """ResnetBlock.forward with temb=None, dropout 0 (model.py:121-141)."""
    h = _conv(P, n + '.conv1', _swish(_gn(P, n + '.norm1', x)), 1)
    h = _conv(P, n + '.conv2', _swish(_gn(P, n + '.norm2', h)), 1)
    if cin != cout:
        x = _conv(P, n + '.nin_shortcut', x, 0)
    return x + h


def _attn(P, n, x):
    """AttnBlock.forward: single head over all channels, scale c^-0.5 (model.py:178-203)."""
    h = _gn(P, n + '.norm', x)
    q, k, v = (_conv(P, n + '.' + t, h, 0) for t in ('q', 'k', 'v'))
    b, c, hh, ww = q.shape
    w = torch.bmm(q.reshape(b, c, -1).permute(0, 2, 1), k.reshape(b, c, -1)) * (int(c) ** (-0.5))
    w = F.softmax(w, dim=2)
    o = torch.bmm(v.reshape(b, c, -1), w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(P, n + '.proj_out', o, 0)


def decode(P, cfg, z, taps=None):
    """decode_first_stage restricted to the AutoencoderKL path: returns images [B, out_ch, 2^(levels-1) R, ...]."""
    x = z / cfg['scale_factor']
    x = _conv(P, 'post_quant_conv', x, 0)
    mods, _ = structure(cfg)
    for m in mods:
        if m[0] == 'conv':
            x = _conv(P, m[1], x, 1)
        elif m[0] == 'res':
            x = _res(P, m[1], x, m[2], m[3])
        elif m[0] == 'attn':
            x = _attn(P, m[1], x)
        elif m[0] == 'up':
            x = _conv(P, m[1] + '.conv', F.interpolate(x, scale_factor=2.0, mode='nearest'), 1)      # model.py:53-57
        if taps is not None:
            taps[m[1]] = x
    return _conv(P, 'decoder.conv_out', _swish(_gn(P, 'decoder.norm_out', x)), 1)

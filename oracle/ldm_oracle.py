"""ORACLE — test infrastructure, not product code (see oracle/edm_oracle.py header).

CPU restatement of the latent-diffusion denoiser used by BASELINE config 5 (Stable Diffusion v1.5):
  * the eps-net `UNetModel` (models/ldm/modules/diffusionmodules/openaimodel.py:413-741: ResBlock :163-275, Downsample :134-160,
    Upsample :91-119; attention.py: CrossAttention :152-193, BasicTransformerBlock :196-215, SpatialTransformer :218-261,
    GEGLU :36-44; util.py:151-171 timestep_embedding) as a functional forward over a flat parameter dict, and
  * the `CFGPrecond` wrapper (models/networks_edm.py:630-759): sigma <-> t by piecewise-linear interpolation over the
    1000 log-alpha knots, classifier-free guidance, D = x - sigma * eps.
Pinned by tests/golden/ref_ldm.npz (outputs of the real reference classes loaded with this file's parameter recipe,
oracle/gen_golden.py).  Paths are relative to /root/reference/diff-solvers-main/.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

CONFIGS = {
    # models/ldm/configs/stable-diffusion/v1-inference.yaml:29-44
    'sd15': dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1), num_res_blocks=2,
                 channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768, img_resolution=64),
    # reduced nets with the same structure (down/up sampling, transformer blocks at two resolutions, head dims 32 and 64)
    'tiny_ldm': dict(in_channels=4, out_channels=4, model_channels=64, attention_resolutions=(2, 1), num_res_blocks=1,
                     channel_mult=(1, 2), num_heads=2, context_dim=128, img_resolution=16),
}


def structure(cfg):
    """Module list of UNetModel.__init__ (openaimodel.py:506-689, use_spatial_transformer=True, legacy=False, depth 1).
    Returns (layers, shapes): layers = [('input_blocks.1', [('res', name, cin, cout), ('attn', name, ch, heads, dim_head), ...]), ...]"""
    mc, mult, nrb, attn = cfg['model_channels'], cfg['channel_mult'], cfg['num_res_blocks'], cfg['attention_resolutions']
    heads = cfg['num_heads']
    inp = [('input_blocks.0', [('conv', 'input_blocks.0.0', cfg['in_channels'], mc)])]
    chans = [mc]
    ch, ds, idx = mc, 1, 1
    for level, m in enumerate(mult):
        for _ in range(nrb):
            layers = [('res', f'input_blocks.{idx}.0', ch, m * mc)]
            ch = m * mc
            if ds in attn:
                layers.append(('attn', f'input_blocks.{idx}.1', ch, heads, ch // heads))
            inp.append((f'input_blocks.{idx}', layers))
            chans.append(ch)
            idx += 1
        if level != len(mult) - 1:
            inp.append((f'input_blocks.{idx}', [('down', f'input_blocks.{idx}.0', ch, ch)]))
            chans.append(ch)
            idx += 1
            ds *= 2
    mid = [('middle_block', [('res', 'middle_block.0', ch, ch), ('attn', 'middle_block.1', ch, heads, ch // heads),
                             ('res', 'middle_block.2', ch, ch)])]
    out = []
    idx = 0
    for level, m in list(enumerate(mult))[::-1]:
        for i in range(nrb + 1):
            ich = chans.pop()
            layers = [('res', f'output_blocks.{idx}.0', ch + ich, mc * m)]
            ch = mc * m
            k = 1
            if ds in attn:
                layers.append(('attn', f'output_blocks.{idx}.{k}', ch, heads, ch // heads))
                k += 1
            if level and i == nrb:
                layers.append(('up', f'output_blocks.{idx}.{k}', ch, ch))
                ds //= 2
            out.append((f'output_blocks.{idx}', layers))
            idx += 1
    return inp, mid, out, ch


def param_shapes(cfg):
    """Ordered (name -> shape) of UNetModel.state_dict() for this config."""
    mc, ctx = cfg['model_channels'], cfg['context_dim']
    ted = mc * 4
    sh = OrderedDict()

    def lin(n, fin, fout, bias=True):
        sh[n + '.weight'] = (fout, fin)
        if bias:
            sh[n + '.bias'] = (fout,)

    def conv(n, cin, cout, k):
        sh[n + '.weight'] = (cout, cin, k, k)
        sh[n + '.bias'] = (cout,)

    def norm(n, c):
        sh[n + '.weight'] = (c,)
        sh[n + '.bias'] = (c,)

    def res(n, cin, cout):
        norm(n + '.in_layers.0', cin)
        conv(n + '.in_layers.2', cin, cout, 3)
        lin(n + '.emb_layers.1', ted, cout)
        norm(n + '.out_layers.0', cout)
        conv(n + '.out_layers.3', cout, cout, 3)
        if cin != cout:
            conv(n + '.skip_connection', cin, cout, 1)

    def attn(n, ch, heads, dh):
        inner = heads * dh
        norm(n + '.norm', ch)
        conv(n + '.proj_in', ch, inner, 1)
        t = n + '.transformer_blocks.0'
        def xattn(a, cdim):
            lin(f'{t}.{a}.to_q', inner, inner, bias=False)
            lin(f'{t}.{a}.to_k', cdim, inner, bias=False)
            lin(f'{t}.{a}.to_v', cdim, inner, bias=False)
            lin(f'{t}.{a}.to_out.0', inner, inner)
        xattn('attn1', inner)                       # registration order of BasicTransformerBlock.__init__ (attention.py:199-206)
        lin(f'{t}.ff.net.0.proj', inner, inner * 8)
        lin(f'{t}.ff.net.2', inner * 4, inner)
        xattn('attn2', ctx)
        for k in (1, 2, 3):
            norm(f'{t}.norm{k}', inner)
        conv(n + '.proj_out', inner, ch, 1)

    lin('time_embed.0', mc, ted)
    lin('time_embed.2', ted, ted)
    inp, mid, out, ch_final = structure(cfg)
    for _, layers in inp + mid + out:
        for L in layers:
            kind, n = L[0], L[1]
            if kind == 'conv':
                conv(n, L[2], L[3], 3)
            elif kind == 'res':
                res(n, L[2], L[3])
            elif kind == 'attn':
                attn(n, L[2], L[3], L[4])
            elif kind == 'down':
                conv(n + '.op', L[2], L[3], 3)
            elif kind == 'up':
                conv(n + '.conv', L[2], L[3], 3)
    norm('out.0', ch_final)
    conv('out.2', mc, cfg['out_channels'], 3)
    return sh


def make_params(name, seed=0):
    """Deterministic parameter recipe (NOT the reference's default init, which zeroes several layers and depends on torch's
    module construction order): every weight ~ U(-1,1) * sqrt(3 / fan_in), biases ~ U(-0.1, 0.1), norm gains 1 + U(-0.1, 0.1).
    The golden generator loads exactly these tensors into the reference classes."""
    cfg = CONFIGS[name]
    g = torch.Generator().manual_seed(seed + 777)
    P = OrderedDict()
    for k, shp in param_shapes(cfg).items():
        if len(shp) == 1:
            is_gain = k.endswith('.weight')
            v = (torch.rand(shp, generator=g) * 2 - 1) * 0.1
            P[k] = v + 1.0 if is_gain else v
        else:
            fan_in = int(np.prod(shp[1:]))
            P[k] = (torch.rand(shp, generator=g) * 2 - 1) * math.sqrt(3.0 / fan_in)
    return P, cfg


# ------------------------------------------------------------------------------------------------ forward

def _gn(P, n, x, eps):
    return F.group_norm(x, 32, P[n + '.weight'], P[n + '.bias'], eps)


def _res(P, n, x, emb):
    """openaimodel.py:255-275 ResBlock._forward (no up/down, no scale-shift norm)."""
    h = F.conv2d(F.silu(_gn(P, n + '.in_layers.0', x, 1e-5)), P[n + '.in_layers.2.weight'], P[n + '.in_layers.2.bias'], padding=1)
    e = F.linear(F.silu(emb), P[n + '.emb_layers.1.weight'], P[n + '.emb_layers.1.bias'])
    h = h + e[:, :, None, None]
    h = F.conv2d(F.silu(_gn(P, n + '.out_layers.0', h, 1e-5)), P[n + '.out_layers.3.weight'], P[n + '.out_layers.3.bias'], padding=1)
    if (n + '.skip_connection.weight') in P:
        x = F.conv2d(x, P[n + '.skip_connection.weight'], P[n + '.skip_connection.bias'])
    return x + h


def _xattn(P, n, x, ctx, heads):
    """attention.py:170-193 CrossAttention.forward (no mask)."""
    q = F.linear(x, P[n + '.to_q.weight'])
    k = F.linear(ctx, P[n + '.to_k.weight'])
    v = F.linear(ctx, P[n + '.to_v.weight'])
    B, L, inner = q.shape
    d = inner // heads
    split = lambda t: t.reshape(B, -1, heads, d).permute(0, 2, 1, 3).reshape(B * heads, -1, d)
    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum('bid,bjd->bij', q, k) * d ** -0.5
    o = torch.einsum('bij,bjd->bid', sim.softmax(dim=-1), v)
    o = o.reshape(B, heads, L, d).permute(0, 2, 1, 3).reshape(B, L, inner)
    return F.linear(o, P[n + '.to_out.0.weight'], P[n + '.to_out.0.bias'])


def _transformer(P, n, x, ctx, heads):
    """attention.py:250-261 SpatialTransformer.forward with one BasicTransformerBlock (:211-215)."""
    B, C, H, W = x.shape
    x_in = x
    x = F.conv2d(_gn(P, n + '.norm', x, 1e-6), P[n + '.proj_in.weight'], P[n + '.proj_in.bias'])
    x = x.permute(0, 2, 3, 1).reshape(B, H * W, -1)
    t = n + '.transformer_blocks.0'
    ln = lambda k, y: F.layer_norm(y, (y.shape[-1],), P[f'{t}.norm{k}.weight'], P[f'{t}.norm{k}.bias'], 1e-5)
    y = ln(1, x)
    x = _xattn(P, t + '.attn1', y, y, heads) + x
    x = _xattn(P, t + '.attn2', ln(2, x), ctx, heads) + x
    h = F.linear(ln(3, x), P[t + '.ff.net.0.proj.weight'], P[t + '.ff.net.0.proj.bias'])
    a, gate = h.chunk(2, dim=-1)                                          # attention.py:42-44 GEGLU
    x = F.linear(a * F.gelu(gate), P[t + '.ff.net.2.weight'], P[t + '.ff.net.2.bias']) + x
    x = x.reshape(B, H, W, -1).permute(0, 3, 1, 2)
    return F.conv2d(x, P[n + '.proj_out.weight'], P[n + '.proj_out.bias']) + x_in


def timestep_embedding(t, dim):
    """util.py:151-171."""
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def unet_forward(P, cfg, x, timesteps, context, taps=None):
    """openaimodel.py:710-741 UNetModel.forward."""
    inp, mid, out, _ = structure(cfg)
    emb = F.linear(timestep_embedding(timesteps, cfg['model_channels']), P['time_embed.0.weight'], P['time_embed.0.bias'])
    emb = F.linear(F.silu(emb), P['time_embed.2.weight'], P['time_embed.2.bias'])

    def run(layers, h):
        for L in layers:
            kind, n = L[0], L[1]
            if kind == 'conv':
                h = F.conv2d(h, P[n + '.weight'], P[n + '.bias'], padding=1)
            elif kind == 'res':
                h = _res(P, n, h, emb)
            elif kind == 'attn':
                h = _transformer(P, n, h, context, L[3])
            elif kind == 'down':
                h = F.conv2d(h, P[n + '.op.weight'], P[n + '.op.bias'], stride=2, padding=1)
            elif kind == 'up':
                h = F.conv2d(F.interpolate(h, scale_factor=2, mode='nearest'), P[n + '.conv.weight'], P[n + '.conv.bias'], padding=1)
            if taps is not None:
                taps[n] = h
        return h
    hs = []
    h = x
    for _, layers in inp:
        h = run(layers, h)
        hs.append(h)
    h = run(mid[0][1], h)
    for _, layers in out:
        h = run(layers, torch.cat([h, hs.pop()], dim=1))
    return F.conv2d(F.silu(_gn(P, 'out.0', h, 1e-5)), P['out.2.weight'], P['out.2.bias'], padding=1)


def make_alphas_cumprod(linear_start=0.00085, linear_end=0.0120, n=1000):
    """models/ldm/modules/diffusionmodules/util.py make_beta_schedule('linear') as configured in v1-inference.yaml:5-6."""
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n, dtype=torch.float64) ** 2
    return torch.cumprod(1.0 - betas, dim=0).to(torch.float32)


class OracleCFGNet:
    """networks_edm.py:630-759 CFGPrecond over the functional eps-net.  net(x, sigma, condition=, unconditional_condition=)."""

    def __init__(self, P, cfg, guidance_rate=7.5, guidance_type='classifier-free', epsilon_t=1e-3):
        self.P, self.cfg = P, cfg
        self.img_resolution, self.img_channels, self.label_dim = cfg['img_resolution'], cfg['in_channels'], True
        self.guidance_rate, self.guidance_type = guidance_rate, guidance_type
        log_alphas = 0.5 * torch.log(make_alphas_cumprod())
        self.M = len(log_alphas)
        self.t_array = torch.linspace(0., 1., self.M + 1)[1:].reshape((1, -1))
        self.log_alpha_array = log_alphas.reshape((1, -1))
        self.sigma_min = float(self.sigma(epsilon_t))
        self.sigma_max = float(self.sigma(1))
        self.taps = None

    @staticmethod
    def interpolate_fn(x, xp, yp):
        """networks_edm.py:711-756 piecewise-linear interpolation through sort/gather."""
        N, K = x.shape[0], xp.shape[1]
        all_x = torch.cat([x.unsqueeze(2), xp.unsqueeze(0).repeat((N, 1, 1))], dim=2)
        sorted_all_x, x_indices = torch.sort(all_x, dim=2)
        x_idx = torch.argmin(x_indices, dim=2)
        cand = x_idx - 1
        start_idx = torch.where(torch.eq(x_idx, 0), torch.tensor(1), torch.where(torch.eq(x_idx, K), torch.tensor(K - 2), cand))
        end_idx = torch.where(torch.eq(start_idx, cand), start_idx + 2, start_idx + 1)
        start_x = torch.gather(sorted_all_x, dim=2, index=start_idx.unsqueeze(2)).squeeze(2)
        end_x = torch.gather(sorted_all_x, dim=2, index=end_idx.unsqueeze(2)).squeeze(2)
        start_idx2 = torch.where(torch.eq(x_idx, 0), torch.tensor(0), torch.where(torch.eq(x_idx, K), torch.tensor(K - 2), cand))
        ypx = yp.unsqueeze(0).expand(N, -1, -1)
        start_y = torch.gather(ypx, dim=2, index=start_idx2.unsqueeze(2)).squeeze(2)
        end_y = torch.gather(ypx, dim=2, index=(start_idx2 + 1).unsqueeze(2)).squeeze(2)
        return start_y + (x - start_x) * (end_y - start_y) / (end_x - start_x)

    def marginal_log_mean_coeff(self, t):
        t = torch.as_tensor(t, dtype=torch.float32)
        return self.interpolate_fn(t.reshape((-1, 1)), self.t_array, self.log_alpha_array).reshape((-1))

    def sigma(self, t):
        lm = self.marginal_log_mean_coeff(t)
        return torch.sqrt(1. - torch.exp(2. * lm)) / torch.exp(lm)

    def sigma_inv(self, sigma):
        lamb = -(sigma.log())
        log_alpha = -0.5 * torch.logaddexp(torch.zeros((1,)), -2. * lamb)
        t = self.interpolate_fn(log_alpha.reshape((-1, 1)), torch.flip(self.log_alpha_array, [1]), torch.flip(self.t_array, [1]))
        return t.reshape((-1,))

    def __call__(self, x, sigma, condition=None, unconditional_condition=None, **_):
        x = x.to(torch.float32)
        sigma = torch.as_tensor(sigma).to(torch.float32).reshape(-1,)
        c_out = -sigma
        c_in = 1 / (sigma ** 2 + 1).sqrt()
        c_noise = self.M * self.sigma_inv(sigma) - 1.
        if c_noise.reshape((-1,)).shape[0] == 1:
            c_noise = c_noise.expand((x.shape[0]))
        eps = lambda xi, ti, ci: unet_forward(self.P, self.cfg, xi, ti, ci, taps=self.taps)
        if self.guidance_rate == 1. or unconditional_condition is None:
            F_x = eps(c_in.reshape(-1, 1, 1, 1) * x, c_noise, condition)
        else:
            x_in = torch.cat([c_in.reshape(-1, 1, 1, 1) * x] * 2)
            t_in = torch.cat([c_noise] * 2)
            u, c = eps(x_in, t_in, torch.cat([unconditional_condition, condition])).chunk(2)
            F_x = u + self.guidance_rate * (c - u)
        return x + c_out.reshape(-1, 1, 1, 1) * F_x

    def round_sigma(self, sigma):
        return torch.as_tensor(sigma)

"""ORACLE — test infrastructure, not product code (see oracle/edm_oracle.py header).

CPU restatement of the AMED path: the predictor network (amed-solver-main/training/networks.py:121-155), the bottleneck
read-out (amed-solver-main/solvers_amed.py:7-55) and the AMED samplers (solvers_amed.py:69-631).
Pinned by tests/golden/ref_amed.npz (real reference outputs, oracle/gen_golden.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .solvers_oracle import dpm_pp_update, dynamic_thresholding, get_schedule


def predictor_forward(W, cfg, bottleneck, t_cur, t_next):
    """networks.py:121-155.  W: state dict; cfg: dict(scale_dir, scale_time)."""
    def lin(n, x):
        y = x @ W[n + '.weight'].t()
        return y + W[n + '.bias'] if (n + '.bias') in W else y

    def temb(t):                                     # :123-126 positional embedding (8 channels, endpoint) with sin/cos swap
        freqs = torch.arange(0, 4, dtype=torch.float32) / (4 - 1)
        e = t.reshape(1,).ger((1 / 10000) ** freqs)
        e = torch.cat([e.cos(), e.sin()], dim=1)
        e = e.reshape(1, 2, -1).flip(1).reshape(1, -1)
        return F.silu(lin('map_layer0', e))
    B = bottleneck.shape[0]
    emb = torch.cat((temb(t_cur).repeat(B, 1), temb(t_next).repeat(B, 1)), dim=1)
    z = lin('enc_layer1', F.silu(lin('enc_layer0', bottleneck.reshape(B, -1))))
    out = torch.cat((z, emb), dim=1)
    r = torch.sigmoid(lin('fc_r', out))
    sd = st = None
    if cfg['scale_dir']:
        sd = torch.sigmoid(lin('fc_scale_dir', out)) / (1 / (2 * cfg['scale_dir'])) + (1 - cfg['scale_dir'])
    if cfg['scale_time']:
        st = torch.sigmoid(lin('fc_scale_time', out)) / (1 / (2 * cfg['scale_time'])) + (1 - cfg['scale_time'])
    one = torch.ones_like(r)
    return [v.reshape(-1, 1, 1, 1) for v in (r, one if sd is None else sd, one if st is None else st)]


@torch.no_grad()
def sample_amed(net, latents, solver, W, cfg, num_steps, afs=False, max_order=None, predict_x0=True, lower_order_final=True,
                sigma_min=0.002, sigma_max=80, schedule_type='polynomial', schedule_rho=7, bottleneck_block=None, class_labels=None,
                condition=None, unconditional_condition=None):
    """solver in {'amed', 'euler', 'ipndm', 'dpm_2', 'dpm_pp'} (solvers_amed.py:69-159, 163-257, 262-396, 400-494, 498-631).
    Bottleneck tap (solvers_amed.py:7-27): EDM nets -> enc['8x8_block2'] with class labels, enc['8x8_block3'] without (:16);
    latent-diffusion nets (guidance_type) -> middle_block output (:12), and under classifier-free guidance the conditional half
    [B:] of the doubled batch (:24-25)."""
    is_ldm = hasattr(net, 'guidance_type')
    t_steps = get_schedule(num_steps, sigma_min, sigma_max, schedule_type=schedule_type, schedule_rho=schedule_rho, net=net)
    B = latents.shape[0]
    if bottleneck_block is None:
        bottleneck_block = 'middle_block.2' if is_ldm else ('enc.8x8_block2' if class_labels is not None else 'enc.8x8_block3')

    def call(x, t):
        if is_ldm:
            return net(x, t, condition=condition, unconditional_condition=unconditional_condition)
        return net(x, t, class_labels=class_labels)
    x_next = latents * t_steps[0]
    hist, hist_t = [], []
    total = 2 * num_steps - 1
    max_order = max_order or (3 if solver == 'dpm_pp' else 4)

    def D_tap(x, t):
        net.taps = {}
        den = call(x, t)
        enc = torch.mean(net.taps[bottleneck_block], dim=1)
        net.taps = None
        if is_ldm and net.guidance_type == 'classifier-free' and enc.shape[0] == 2 * B:
            enc = enc[B:]
        return den, enc

    def push(d):
        if len(hist) == max_order - 1:
            for k in range(max_order - 2):
                hist[k] = hist[k + 1]
            hist[-1] = d
        else:
            hist.append(d)

    def ab(d, order):
        if order == 1:
            return d
        if order == 2:
            return (3 * d - hist[-1]) / 2
        if order == 3:
            return (23 * d - 16 * hist[-1] + 5 * hist[-2]) / 12
        return (55 * d - 59 * hist[-1] + 37 * hist[-2] - 9 * hist[-3]) / 24

    for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
        x_cur = x_next
        use_afs = afs and (i == 0 if solver in ('amed', 'euler', 'dpm_2') else len(hist) == 0)
        if use_afs:
            d_cur = x_cur / ((1 + t_cur ** 2).sqrt())
            den = x_cur - t_cur * d_cur
            enc = torch.zeros((B, 8, 8))
        else:
            den, enc = D_tap(x_cur, t_cur)
            d_cur = (x_cur - den) / t_cur
        tc, tn = t_cur.reshape(-1, 1, 1, 1), t_next.reshape(-1, 1, 1, 1)
        r, sd, st = predictor_forward(W, cfg, enc, tc, tn)
        t_mid = (tn ** r) * (tc ** (1 - r))
        if solver in ('amed', 'euler', 'dpm_2'):
            x_mid = x_cur + (t_mid - tc) * d_cur
            d_mid = (x_mid - call(x_mid, st * t_mid)) / t_mid
            if solver == 'amed':
                x_next = x_cur + sd * (tn - tc) * d_mid
            elif solver == 'euler':
                x_next = x_mid + sd * (tn - t_mid) * d_mid
            else:
                x_next = x_cur + sd * (tn - tc) * ((1 / (2 * r)) * d_mid + (1 - 1 / (2 * r)) * d_cur)
        elif solver == 'ipndm':
            order = min(max_order, len(hist) + 1)
            x_mid = x_cur + (t_mid - tc) * ab(d_cur, order)
            push(d_cur)
            order = min(max_order, len(hist) + 1)
            d_mid = (x_mid - call(x_mid, st * t_mid)) / t_mid
            x_next = x_mid + sd * (tn - t_mid) * ab(d_mid, order)
            push(d_mid)
        elif solver == 'dpm_pp':
            step = 2 * i + 1
            hist.append(dynamic_thresholding(den) if predict_x0 else d_cur)
            hist_t.append(tc)
            order = (step if step < max_order else min(max_order, total - step)) if lower_order_final else min(max_order, step)
            x_mid = dpm_pp_update(x_cur, hist, hist_t, t_mid, order, predict_x0=predict_x0)
            step += 1
            den2 = call(x_mid, st * t_mid)
            hist.append(dynamic_thresholding(den2) if predict_x0 else (x_mid - den2) / t_mid)
            hist_t.append(t_mid)
            order = (step if step < max_order else min(max_order, total - step)) if lower_order_final else min(step, max_order)
            x_next = dpm_pp_update(x_mid, hist, hist_t, tn, order, predict_x0=predict_x0, scale=sd)
            hist, hist_t = hist[-3:], hist_t[-3:]
        else:
            raise NotImplementedError(solver)
    return x_next

"""Plan compiler for the latent-diffusion eps-net (Stable Diffusion v1.x `UNetModel`, BASELINE config 5).

Reference being lowered: models/ldm/modules/diffusionmodules/openaimodel.py:710-741 (UNetModel.forward), ResBlock :255-275,
Downsample :134-160, Upsample :91-119; models/ldm/modules/attention.py SpatialTransformer :250-261, BasicTransformerBlock :211-215,
CrossAttention :170-193, GEGLU :42-44; util.py:151-171 (timestep_embedding).  Same op set and executor as the EDM nets (plan.py):
every contraction is the tcgen05 GEMM kernel; LayerNorm / GEGLU / softmax / GroupNorm are the HBM-bound companions.

Layout notes specific to this net:
  * head dims 40 / 80 / 160 are zero-padded to 64 / 128 / 192 inside the packed q/k/v/out weights (K blocks are 64 wide);
  * the 77 context tokens are not padded in memory: K extents / key counts that are not multiples of 64 are zero-filled by TMA;
  * the stride-2 Downsample conv runs on a space-to-depth repack (gn_apply resample=3) with a per-tap (shift, phase) table;
  * classifier-free guidance evaluates the batch [uncond | cond] = 2B samples in one pass; eps is written NCHW.
"""
import math
import re
from collections import OrderedDict

import torch

from . import _cstructs as S
from . import gemm_desc as G
from .plan import Plan, WeightBlob, _Arena

F4, H2 = 4, 2
CTX_TOKENS_PITCH = 128        # P / V^T row pitch for the 77 context tokens (multiple of 8 elements for TMA strides)


def _groups(c):
    """LDM uses GroupNorm32(32, channels) everywhere (util.py:202-216, attention.py:76-77): always 32 groups."""
    assert c % 32 == 0
    return 32


def dpad(d):
    return -(-d // 64) * 64


def prows(n):
    """Row count of a weight packed by gemm_desc.pack_conv_weight (padded to whole N tiles)."""
    bn, tiles = G.pick_bn(n)
    return bn * tiles


def ldm_structure(params, num_heads):
    """Block structure from UNetModel.state_dict() names/shapes:
    [(block_name, [('conv'|'res'|'attn'|'down'|'up', module_name, ...)])] for input / middle / output blocks."""
    names = list(params.keys())
    mods = OrderedDict()
    for k in names:
        m = re.match(r'((?:input_blocks|output_blocks)\.\d+\.\d+|middle_block\.\d+)\.', k)
        if m:
            mods.setdefault(m.group(1), []).append(k)

    def kind_of(mod, keys):
        if any('.in_layers.' in k for k in keys):
            w = params[mod + '.in_layers.2.weight']
            return ('res', mod, w.shape[1], w.shape[0])
        if any('.transformer_blocks.' in k for k in keys):
            ch = params[mod + '.norm.weight'].shape[0]
            inner = params[mod + '.proj_in.weight'].shape[0]
            return ('attn', mod, ch, num_heads, inner // num_heads)
        if (mod + '.op.weight') in params:
            w = params[mod + '.op.weight']
            return ('down', mod, w.shape[1], w.shape[0])
        if (mod + '.conv.weight') in params:
            w = params[mod + '.conv.weight']
            return ('up', mod, w.shape[1], w.shape[0])
        w = params[mod + '.weight']
        return ('conv', mod, w.shape[1], w.shape[0])

    def block_key(mod):
        parts = mod.split('.')
        return parts[0] if parts[0] == 'middle_block' else parts[0] + '.' + parts[1]
    blocks = OrderedDict()
    for mod, keys in mods.items():
        blocks.setdefault(block_key(mod), []).append(kind_of(mod, keys))
    inp = [(b, l) for b, l in blocks.items() if b.startswith('input_blocks')]
    mid = [(b, l) for b, l in blocks.items() if b.startswith('middle_block')]
    out = [(b, l) for b, l in blocks.items() if b.startswith('output_blocks')]
    inp.sort(key=lambda t: int(t[0].split('.')[1]))
    out.sort(key=lambda t: int(t[0].split('.')[1]))
    return dict(inp=inp, mid=mid, out=out, model_channels=params['time_embed.0.weight'].shape[1], ted=params['time_embed.0.weight'].shape[0],
                in_channels=params['input_blocks.0.0.weight'].shape[1], out_channels=params['out.2.weight'].shape[0], num_heads=num_heads)


def _pad_heads_rows(w, heads, dh):
    """[heads*dh, K] -> [heads*dpad, K] with zero rows."""
    dp = dpad(dh)
    out = torch.zeros(heads * dp, w.shape[1])
    out.view(heads, dp, -1)[:, :dh] = w.reshape(heads, dh, -1)
    return out


def _pad_heads_cols(w, heads, dh):
    """[N, heads*dh] -> [N, heads*dpad] with zero columns."""
    dp = dpad(dh)
    out = torch.zeros(w.shape[0], heads * dp)
    out.view(w.shape[0], heads, dp)[:, :, :dh] = w.reshape(w.shape[0], heads, dh)
    return out


def pack_ldm_weights(st, params, f8=False, f8_linear=False):
    """f8=True: the ResBlock convolutions (in_layers.2, out_layers.3 + skip_connection) are packed for the f8 GEMM mode (csrc/ops.h).
    f8_linear=True (opt-in, needs f8): also the transformer linears whose A operand has a single consumer -- proj_in, attn2.to_q, the
    GEGLU feed-forward pair and proj_out (75 % of the transformer's linear FLOPs)."""
    assert f8 or not f8_linear
    P = lambda k: params[k].detach().float().cpu()
    wb = WeightBlob()
    info = dict(res=[], ctx_dim=None, f8_shift={})

    def add_lin(key, w, bias=None, as_f8=False):
        """Linear / 1x1 conv weight [N, K] as a packed GEMM operand."""
        if as_f8:
            packed, info['f8_shift'][key] = G.pack_conv_weight_f8(w.reshape(w.shape[0], w.shape[1], 1, 1))
            wb.add(key + ':w', packed)
        else:
            wb.add(key + ':w', G.pack_conv_weight(w.reshape(w.shape[0], w.shape[1], 1, 1)))
        if bias is not None:
            wb.add(key + ':b', bias)

    def add_conv(key, w, skip_w=None, bias=None, as_f8=False):
        if as_f8:
            packed, info['f8_shift'][key] = G.pack_conv_weight_f8(w, skip_w)
            wb.add(key + ':w', packed)
        else:
            wb.add(key + ':w', G.pack_conv_weight(w, skip_w))
        if bias is not None:
            wb.add(key + ':b', bias)

    for k in ('time_embed.0', 'time_embed.2'):
        wb.add(k + ':w', P(k + '.weight'))
        wb.add(k + ':b', P(k + '.bias'))
    aff_w, aff_b = [], []
    aff_off = 0
    for _, layers in st['inp'] + st['mid'] + st['out']:
        for L in layers:
            kind, n = L[0], L[1]
            if kind == 'conv':
                add_conv(n, P(n + '.weight'), bias=P(n + '.bias'))
            elif kind == 'res':
                wb.add(n + '.n0:g', P(n + '.in_layers.0.weight'))
                wb.add(n + '.n0:b', P(n + '.in_layers.0.bias'))
                add_conv(n + '.c0', P(n + '.in_layers.2.weight'), bias=P(n + '.in_layers.2.bias'), as_f8=f8)
                wb.add(n + '.n1:g', P(n + '.out_layers.0.weight'))
                wb.add(n + '.n1:b', P(n + '.out_layers.0.bias'))
                b1 = P(n + '.out_layers.3.bias')
                skw = None
                if (n + '.skip_connection.weight') in params:
                    skw = P(n + '.skip_connection.weight')
                    b1 = b1 + P(n + '.skip_connection.bias')
                add_conv(n + '.c1', P(n + '.out_layers.3.weight'), skw, bias=b1, as_f8=f8)
                aff_w.append(P(n + '.emb_layers.1.weight'))
                aff_b.append(P(n + '.emb_layers.1.bias'))
                info['res'].append((n, aff_off))
                aff_off += aff_w[-1].shape[0]
            elif kind == 'attn':
                _, _, ch, heads, dh = L
                t = n + '.transformer_blocks.0'
                wb.add(n + '.norm:g', P(n + '.norm.weight'))
                wb.add(n + '.norm:b', P(n + '.norm.bias'))
                add_lin(n + '.proj_in', P(n + '.proj_in.weight').reshape(heads * dh, ch), P(n + '.proj_in.bias'), as_f8=f8_linear)
                for k in (1, 2, 3):
                    wb.add(f'{t}.norm{k}:g', P(f'{t}.norm{k}.weight'))
                    wb.add(f'{t}.norm{k}:b', P(f'{t}.norm{k}.bias'))
                # self-attention: [q | k] rows for one GEMM, v as the M operand of the V^T GEMM
                wq, wk, wv = (_pad_heads_rows(P(f'{t}.attn1.to_{x}.weight'), heads, dh) for x in 'qkv')
                add_lin(t + '.attn1.qk', torch.cat([wq, wk]))
                wb.add(t + '.attn1.v:w', G.split_planes(wv))
                add_lin(t + '.attn1.out', _pad_heads_cols(P(t + '.attn1.to_out.0.weight'), heads, dh), P(t + '.attn1.to_out.0.bias'))
                # cross-attention
                add_lin(t + '.attn2.q', _pad_heads_rows(P(t + '.attn2.to_q.weight'), heads, dh), as_f8=f8_linear)
                add_lin(t + '.attn2.k', _pad_heads_rows(P(t + '.attn2.to_k.weight'), heads, dh))
                wb.add(t + '.attn2.v:w', G.split_planes(_pad_heads_rows(P(t + '.attn2.to_v.weight'), heads, dh)))
                add_lin(t + '.attn2.out', _pad_heads_cols(P(t + '.attn2.to_out.0.weight'), heads, dh), P(t + '.attn2.to_out.0.bias'))
                info['ctx_dim'] = P(t + '.attn2.to_k.weight').shape[1]
                add_lin(t + '.ff1', P(t + '.ff.net.0.proj.weight'), P(t + '.ff.net.0.proj.bias'), as_f8=f8_linear)
                add_lin(t + '.ff2', P(t + '.ff.net.2.weight'), P(t + '.ff.net.2.bias'), as_f8=f8_linear)
                add_lin(n + '.proj_out', P(n + '.proj_out.weight').reshape(ch, heads * dh), P(n + '.proj_out.bias'), as_f8=f8_linear)
            elif kind == 'down':
                add_conv(n, P(n + '.op.weight'), bias=P(n + '.op.bias'))
            elif kind == 'up':
                add_conv(n, P(n + '.conv.weight'), bias=P(n + '.conv.bias'), as_f8=f8)
    wb.add('affine:w', torch.cat(aff_w, dim=0))
    wb.add('affine:b', torch.cat(aff_b, dim=0))
    info['aff_total'] = aff_off
    wb.add('out.0:g', P('out.0.weight'))
    wb.add('out.0:b', P('out.0.bias'))
    add_conv('out.2', P('out.2.weight'), bias=P('out.2.bias'))
    return wb, info


def compile_ldm_plan(st, wb, info, B, Bt, nT, R, npass=3, ctx_tokens=77, flash_attn=True, f8=False, f8_linear=False):
    """Lower the eps-net for Bt samples (Bt = B, or 2B under classifier-free guidance) at latent resolution R.
    nT in {1, Bt}: number of timestep values.  io: X = x [B,C,R,R], SIGMA = timesteps [nT], LABELS = coef [B|1][4] (c_in in slot 2),
    CTX = context [Bt, 77, ctx_dim], D = eps [Bt,C,R,R] (NCHW), BOTTLENECK = channel-mean of the middle block [Bt, 64]."""
    assert nT in (1, Bt)
    assert not f8 or (npass == 3 and info['f8_shift'])
    assert f8 or not f8_linear
    fmt_res = 1 if f8 else 0
    fmt_lin = 1 if f8_linear else 0

    def f8_args(key):
        return dict(f8=True, acc_scale=2.0 ** -info['f8_shift'][key]) if key in info['f8_shift'] else {}
    A = _Arena()
    ops = []
    npl = 2
    io = lambda slot: S.ref(S.SPACE_IO, slot)
    W = wb.ref
    mc, ted = st['model_channels'], st['ted']
    cd = info['ctx_dim']
    T = ctx_tokens
    TP = CTX_TOKENS_PITCH
    tag = [0]
    emit = lambda b: ops.append((tag[0], b))
    n_gn = sum(2 if L[0] == 'res' else (1 if L[0] == 'attn' else 0) for _, ls in st['inp'] + st['mid'] + st['out'] for L in ls) + 1
    A.need('stats', n_gn * Bt * 32 * 2 * 8)
    stat_i = [0]

    def stats_slot():
        i = stat_i[0]
        stat_i[0] += 1
        return i * Bt * 32 * 2 * 8

    emit(lambda R_: S.MemsetDesc(ptr=R_('stats'), bytes=n_gn * Bt * 32 * 2 * 8))
    # ---------------- timestep embedding (util.py:151-171, openaimodel.py:723-724) + all emb_layers in one launch -------------
    A.need('emb0', nT * mc * F4)
    A.need('e1', nT * ted * F4)
    A.need('e2', nT * ted * F4)
    A.need('aff', nT * info['aff_total'] * F4)
    emit(lambda R_: S.PosembDesc(sigma=io(S.DS_IO_SIGMA), nsig=nT, num_channels=mc, endpoint=0, swap_sincos=0, sigma_data=0.5, mode=1,
                                 coef=0, emb=R_('emb0')))
    emit(lambda R_: S.LinearDesc(in_=R_('emb0'), in_stride=mc if nT > 1 else 0, W=W('time_embed.0:w'), b=W('time_embed.0:b'), out=R_('e1'),
                                 n_rows=nT, in_f=mc, out_f=ted, act=1, in_scale=1.0))
    # every consumer applies SiLU first (ResBlock.emb_layers = SiLU -> Linear, openaimodel.py:205-211): store silu(emb)
    emit(lambda R_: S.LinearDesc(in_=R_('e1'), in_stride=ted if nT > 1 else 0, W=W('time_embed.2:w'), b=W('time_embed.2:b'), out=R_('e2'),
                                 n_rows=nT, in_f=ted, out_f=ted, act=1, in_scale=1.0))
    emit(lambda R_: S.LinearDesc(in_=R_('e2'), in_stride=ted if nT > 1 else 0, W=W('affine:w'), b=W('affine:b'), out=R_('aff'), n_rows=nT,
                                 in_f=ted, out_f=info['aff_total'], act=0, in_scale=1.0))
    aff_stride = info['aff_total'] if nT > 1 else 0
    aff_off = dict(info['res'])
    # ---------------- context tokens -> fp16 planes (once per forward, shared by every cross-attention) ------------------------
    A.need('ctx', npl * Bt * T * cd * H2)
    emit(lambda R_: S.GnApplyDesc(src0=io(S.DS_IO_CTX), src1=0, C0=cd, C1=0, H=T, W=1, B=Bt, groups=32, sums=0, gamma=0, beta=0, eps=0.0,
                                  silu=0, ada=0, ada_stride=0, resample=0, nplanes=npl, out_act=0, out_raw=R_('ctx'), out_raw_f32=0))

    def gn_stats(slot, parts, hw):
        (n0, c0), (n1, c1) = parts[0], (parts[1] if len(parts) > 1 else (None, 0))
        emit(lambda R_: S.GnStatsDesc(src0=R_(n0), src1=R_(n1) if n1 else 0, C0=c0, C1=c1, HW=hw, B=Bt, groups=_groups(c0 + c1),
                                      sums=R_('stats', slot)))

    def gn_apply(slot, parts, H, g, b, eps, silu, out, fmt=0):
        (n0, c0), (n1, c1) = parts[0], (parts[1] if len(parts) > 1 else (None, 0))
        emit(lambda R_: S.GnApplyDesc(src0=R_(n0), src1=R_(n1) if n1 else 0, C0=c0, C1=c1, H=H, W=H, B=Bt, groups=_groups(c0 + c1),
                                      sums=R_('stats', slot), gamma=W(g), beta=W(b), eps=eps, silu=silu, ada=0, ada_stride=0, resample=0,
                                      nplanes=npl, out_act=R_(out), out_raw=0, out_raw_f32=0, fmt=fmt))

    def lower_res(L, parts, H):
        """ResBlock (openaimodel.py:255-275): GN+SiLU+conv3x3, + Linear(SiLU(emb)), GN+SiLU+conv3x3, + skip (identity | 1x1)."""
        _, n, cin, cout = L
        M = Bt * H * H
        assert sum(c for _, c in parts) == cin
        s0 = stats_slot()
        gn_stats(s0, parts, H * H)
        A.need('act', npl * M * max(cin, cout) * H2)
        has_skip = (n + '.c1:w') in wb.off and cin != cout
        if has_skip:
            A.need('raw', npl * M * cin * H2)
        (n0, c0), (n1, c1) = parts[0], (parts[1] if len(parts) > 1 else (None, 0))
        emit(lambda R_: S.GnApplyDesc(src0=R_(n0), src1=R_(n1) if n1 else 0, C0=c0, C1=c1, H=H, W=H, B=Bt, groups=_groups(cin),
                                      sums=R_('stats', s0), gamma=W(n + '.n0:g'), beta=W(n + '.n0:b'), eps=1e-5, silu=1, ada=0, ada_stride=0,
                                      resample=0, nplanes=npl, out_act=R_('act'), out_raw=R_('raw') if has_skip else 0, out_raw_f32=0,
                                      fmt=fmt_res))
        A.need('y', M * cout * F4)
        off = aff_off[n]
        emit(lambda R_: G.conv_gemm(R_('act'), Bt, H, H, cin, W(n + '.c0:w'), cout, taps=9, npass=npass, out_f32=R_('y'), bias=W(n + '.c0:b'),
                                    rowvec=R_('aff', off * F4), rowvec_stride=aff_stride, **f8_args(n + '.c0'))[0])
        s1 = stats_slot()
        gn_stats(s1, [('y', cout)], H * H)
        gn_apply(s1, [('y', cout)], H, n + '.n1:g', n + '.n1:b', 1e-5, 1, 'act', fmt=fmt_res)
        out = A.need('h:' + n, M * cout * F4)
        res_name = None if has_skip else parts[0][0]
        assert has_skip or len(parts) == 1
        emit(lambda R_: G.conv_gemm(R_('act'), Bt, H, H, cout, W(n + '.c1:w'), cout, taps=9, npass=npass, a2_ptr=R_('raw') if has_skip else 0,
                                    C2=cin if has_skip else 0, out_f32=R_(out), bias=W(n + '.c1:b'), residual=R_(res_name) if res_name else 0,
                                    ldr=cout, scale=1.0, **f8_args(n + '.c1'))[0])
        return out, cout

    def cast_planes(src, C, H, dst, fmt=0):
        """fp32 NHWC -> fp16 hi/lo planes, or the f8 operand image (fmt=1); no normalisation."""
        emit(lambda R_: S.GnApplyDesc(src0=R_(src), src1=0, C0=C, C1=0, H=H, W=H, B=Bt, groups=32, sums=0, gamma=0, beta=0, eps=0.0, silu=0,
                                      ada=0, ada_stride=0, resample=0, nplanes=npl, out_act=0, out_raw=R_(dst), out_raw_f32=0, fmt=fmt))

    def lower_attn(L, src, H):
        """SpatialTransformer with one BasicTransformerBlock (attention.py:250-261, :211-215)."""
        _, n, ch, heads, dh = L
        t = n + '.transformer_blocks.0'
        inner, dp = heads * dh, dpad(dh)
        hp = heads * dp
        Lq = H * H
        M = Bt * Lq
        s0 = stats_slot()
        gn_stats(s0, [(src, ch)], Lq)
        A.need('act', npl * M * max(ch, inner) * H2)
        gn_apply(s0, [(src, ch)], H, n + '.norm:g', n + '.norm:b', 1e-6, 0, 'act', fmt=fmt_lin)
        for nm in ('t0', 't1', 't2', 't3'):
            A.need(nm, M * inner * F4)
        emit(lambda R_: G.conv_gemm(R_('act'), Bt, H, H, ch, W(n + '.proj_in:w'), inner, taps=1, npass=npass, out_f32=R_('t0'),
                                    bias=W(n + '.proj_in:b'), **(f8_args(n + '.proj_in') if f8_linear else {}))[0])
        A.need('ln', npl * M * inner * H2)
        A.need('qk', npl * M * 2 * hp * H2)
        A.need('vt', npl * Bt * hp * max(Lq, TP) * H2)
        if not (flash_attn and dp == 64 and npl == 2):
            A.need('S', Bt * heads * Lq * max(Lq, 80) * F4)
            A.need('P', npl * Bt * heads * Lq * max(Lq, TP) * H2)
        A.need('o', npl * M * hp * H2)

        def ln(k, srcbuf, fmt=0):
            emit(lambda R_: S.LayernormDesc(src=R_(srcbuf), gamma=W(f'{t}.norm{k}:g'), beta=W(f'{t}.norm{k}:b'), out=R_('ln'), rows=M, C=inner,
                                            nplanes=npl, eps=1e-5, fmt=fmt))
        # ---- self-attention (attn1): x = attn1(norm1(x)) + x
        ln(1, 't0')
        emit(lambda R_: G.conv_gemm(R_('ln'), Bt, H, H, inner, W(t + '.attn1.qk:w'), 2 * hp, taps=1, npass=npass, out_h16=R_('qk'))[0])
        emit(lambda R_: G.rows_gemm(W(t + '.attn1.v:w'), hp, inner, 1, R_('ln'), Lq, inner, Bt, inner, num_z=Bt, nh=1, m_valid=hp, n_valid=Lq,
                                    npass=npass, b_z_per_zb=1, out_h16=R_('vt'), o_zb=hp * Lq, ldo=Lq, o_plane=Bt * hp * Lq)[0])
        flash = flash_attn and dp == 64 and npl == 2
        if flash:
            # fused QK^T -> softmax -> PV (attention.cu): at 64x64 latents the 4096 x 4096 score matrix per head never reaches HBM
            emit(lambda R_: S.AttnDesc(q=R_('qk'), k=R_('qk'), vt=R_('vt'), out=R_('o'), B=Bt, nh=heads, L=Lq, Lk=Lq, q_pitch=2 * hp, q_c0=0,
                                       k_pitch=2 * hp, k_c0=hp, vt_pitch=Lq, o_pitch=hp, nplanes=npl, scale=dh ** -0.5))
        else:
            emit(lambda R_: G.rows_gemm(R_('qk'), Lq, 2 * hp, Bt, R_('qk'), Lq, 2 * hp, Bt, dp, num_z=Bt * heads, nh=heads, m_valid=Lq, n_valid=Lq,
                                        npass=npass, a_c_per_zh=dp, a_n_per_zb=1, b_k0=hp, b_k_per_zh=dp, b_z_per_zb=1, out_f32=R_('S'),
                                        o_zb=heads * Lq * Lq, o_zh=Lq * Lq, ldo=Lq, scale=dh ** -0.5)[0])
            emit(lambda R_: S.SoftmaxDesc(S=R_('S'), P=R_('P'), rows=Bt * heads * Lq, L=Lq, nplanes=npl, pitch_in=0, pitch_out=0))
            emit(lambda R_: G.rows_gemm(R_('P'), Lq, Lq, Bt * heads, R_('vt'), hp, Lq, Bt, Lq, num_z=Bt * heads, nh=heads, m_valid=Lq, n_valid=dp,
                                        npass=npass, a_n_per_zb=heads, a_n_per_zh=1, b_row_per_zh=dp, b_z_per_zb=1, out_h16=R_('o'),
                                        o_zb=Lq * hp, o_zh=dp, ldo=hp, o_plane=M * hp)[0])
        emit(lambda R_: G.conv_gemm(R_('o'), Bt, H, H, hp, W(t + '.attn1.out:w'), inner, taps=1, npass=npass, out_f32=R_('t1'),
                                    bias=W(t + '.attn1.out:b'), residual=R_('t0'), ldr=inner)[0])
        # ---- cross-attention (attn2): x = attn2(norm2(x), context) + x
        ln(2, 't1', fmt=fmt_lin)           # single consumer: the to_q GEMM below
        A.need('q2', npl * M * hp * H2)
        A.need('k2', npl * Bt * T * hp * H2)
        emit(lambda R_: G.conv_gemm(R_('ln'), Bt, H, H, inner, W(t + '.attn2.q:w'), hp, taps=1, npass=npass, out_h16=R_('q2'),
                                    **(f8_args(t + '.attn2.q') if f8_linear else {}))[0])
        emit(lambda R_: G.rows_gemm(R_('ctx'), Bt * T, cd, 1, W(t + '.attn2.k:w'), prows(hp), cd, 1, cd, num_z=1, nh=1, m_valid=Bt * T, n_valid=hp,
                                    npass=npass, out_h16=R_('k2'), ldo=hp, o_plane=Bt * T * hp)[0])
        emit(lambda R_: G.rows_gemm(W(t + '.attn2.v:w'), hp, cd, 1, R_('ctx'), T, cd, Bt, cd, num_z=Bt, nh=1, m_valid=hp, n_valid=T,
                                    npass=npass, b_z_per_zb=1, out_h16=R_('vt'), o_zb=hp * TP, ldo=TP, o_plane=Bt * hp * TP)[0])
        if flash:
            emit(lambda R_: S.AttnDesc(q=R_('q2'), k=R_('k2'), vt=R_('vt'), out=R_('o'), B=Bt, nh=heads, L=Lq, Lk=T, q_pitch=hp, q_c0=0,
                                       k_pitch=hp, k_c0=0, vt_pitch=TP, o_pitch=hp, nplanes=npl, scale=dh ** -0.5))
        else:
            emit(lambda R_: G.rows_gemm(R_('q2'), Lq, hp, Bt, R_('k2'), T, hp, Bt, dp, num_z=Bt * heads, nh=heads, m_valid=Lq, n_valid=T,
                                        npass=npass, a_c_per_zh=dp, a_n_per_zb=1, b_k_per_zh=dp, b_z_per_zb=1, out_f32=R_('S'),
                                        o_zb=heads * Lq * 80, o_zh=Lq * 80, ldo=80, scale=dh ** -0.5)[0])
            emit(lambda R_: S.SoftmaxDesc(S=R_('S'), P=R_('P'), rows=Bt * heads * Lq, L=T, nplanes=npl, pitch_in=80, pitch_out=TP))
            emit(lambda R_: G.rows_gemm(R_('P'), Lq, TP, Bt * heads, R_('vt'), hp, TP, Bt, TP, num_z=Bt * heads, nh=heads, m_valid=Lq, n_valid=dp,
                                        npass=npass, a_n_per_zb=heads, a_n_per_zh=1, b_row_per_zh=dp, b_z_per_zb=1, out_h16=R_('o'),
                                        o_zb=Lq * hp, o_zh=dp, ldo=hp, o_plane=M * hp, a_k_valid=T, b_k_valid=T)[0])
        emit(lambda R_: G.conv_gemm(R_('o'), Bt, H, H, hp, W(t + '.attn2.out:w'), inner, taps=1, npass=npass, out_f32=R_('t2'),
                                    bias=W(t + '.attn2.out:b'), residual=R_('t1'), ldr=inner)[0])
        # ---- GEGLU feed-forward: x = ff(norm3(x)) + x
        ln(3, 't2', fmt=fmt_lin)
        A.need('ff', M * 8 * inner * F4)
        A.need('gg', npl * M * 4 * inner * H2)
        emit(lambda R_: G.conv_gemm(R_('ln'), Bt, H, H, inner, W(t + '.ff1:w'), 8 * inner, taps=1, npass=npass, out_f32=R_('ff'),
                                    bias=W(t + '.ff1:b'), **(f8_args(t + '.ff1') if f8_linear else {}))[0])
        emit(lambda R_: S.GegluDesc(src=R_('ff'), out=R_('gg'), rows=M, I=4 * inner, nplanes=npl, fmt=fmt_lin))
        emit(lambda R_: G.conv_gemm(R_('gg'), Bt, H, H, 4 * inner, W(t + '.ff2:w'), inner, taps=1, npass=npass, out_f32=R_('t3'),
                                    bias=W(t + '.ff2:b'), residual=R_('t2'), ldr=inner, **(f8_args(t + '.ff2') if f8_linear else {}))[0])
        # ---- proj_out + outer residual
        cast_planes('t3', inner, H, 'ln', fmt=fmt_lin)
        out = A.need('h:' + n, M * ch * F4)
        emit(lambda R_: G.conv_gemm(R_('ln'), Bt, H, H, inner, W(n + '.proj_out:w'), ch, taps=1, npass=npass, out_f32=R_(out),
                                    bias=W(n + '.proj_out:b'), residual=R_(src), ldr=ch, **(f8_args(n + '.proj_out') if f8_linear else {}))[0])
        return out, ch

    def lower_down(L, src, H):
        _, n, cin, cout = L
        Ho = H // 2
        A.need('s2d', npl * Bt * H * H * cin * H2)
        emit(lambda R_: S.GnApplyDesc(src0=R_(src), src1=0, C0=cin, C1=0, H=H, W=H, B=Bt, groups=32, sums=0, gamma=0, beta=0, eps=0.0, silu=0,
                                      ada=0, ada_stride=0, resample=3, nplanes=npl, out_act=0, out_raw=R_('s2d'), out_raw_f32=0))
        out = A.need('h:' + n, Bt * Ho * Ho * cout * F4)
        emit(lambda R_: G.conv_gemm(R_('s2d'), Bt, Ho, Ho, cin, W(n + ':w'), cout, taps=9, npass=npass, out_f32=R_(out), bias=W(n + ':b'),
                                    s2d=True)[0])
        return out, cout, Ho

    def lower_up(L, src, H):
        _, n, cin, cout = L
        Ho = H * 2
        A.need('act', npl * Bt * Ho * Ho * cin * H2)
        emit(lambda R_: S.GnApplyDesc(src0=R_(src), src1=0, C0=cin, C1=0, H=H, W=H, B=Bt, groups=32, sums=0, gamma=0, beta=0, eps=0.0, silu=0,
                                      ada=0, ada_stride=0, resample=2, nplanes=npl, out_act=0, out_raw=R_('act'), out_raw_f32=0, fmt=fmt_res))
        out = A.need('h:' + n, Bt * Ho * Ho * cout * F4)
        emit(lambda R_: G.conv_gemm(R_('act'), Bt, Ho, Ho, cin, W(n + ':w'), cout, taps=9, npass=npass, out_f32=R_(out), bias=W(n + ':b'),
                                    **f8_args(n))[0])
        return out, cout, Ho

    # ---------------- input conv --------------------------------------------------------------------------------------------
    cimg = st['in_channels']
    A.need('in_planes', npl * Bt * R * R * 64 * H2)
    emit(lambda R_: S.PrepInputDesc(x=io(S.DS_IO_X), coef=io(S.DS_IO_LABELS), coef_stride=4 if B > 1 and nT > 1 else 0, B=Bt, C=cimg,
                                    HW=R * R, nplanes=npl, x_batch=B, out=R_('in_planes')))
    first = st['inp'][0][1][0]
    h = A.need('h:' + first[1], Bt * R * R * first[3] * F4)
    emit(lambda R_: G.conv_gemm(R_('in_planes'), Bt, R, R, 64, W(first[1] + ':w'), first[3], taps=9, npass=npass, out_f32=R_(h),
                                bias=W(first[1] + ':b'))[0])
    cur, cur_c, H = h, first[3], R
    hs = [(cur, cur_c)]
    for _, layers in st['inp'][1:]:
        for L in layers:
            tag[0] += 1
            if L[0] == 'res':
                cur, cur_c = lower_res(L, [(cur, cur_c)], H)
            elif L[0] == 'attn':
                cur, cur_c = lower_attn(L, cur, H)
            elif L[0] == 'down':
                cur, cur_c, H = lower_down(L, cur, H)
        hs.append((cur, cur_c))
    for L in st['mid'][0][1]:
        tag[0] += 1
        if L[0] == 'res':
            cur, cur_c = lower_res(L, [(cur, cur_c)], H)
        else:
            cur, cur_c = lower_attn(L, cur, H)
    mid_out, mid_c, mid_H = cur, cur_c, H
    emit(lambda R_: S.ChanmeanDesc(src=R_(mid_out), out=io(S.DS_IO_BOTTLENECK), rows=Bt * mid_H * mid_H, C=mid_c))
    for _, layers in st['out']:
        sk, sc = hs.pop()
        first_layer = True
        for L in layers:
            tag[0] += 1
            if L[0] == 'res':
                parts = [(cur, cur_c), (sk, sc)] if first_layer else [(cur, cur_c)]
                cur, cur_c = lower_res(L, parts, H)
            elif L[0] == 'attn':
                cur, cur_c = lower_attn(L, cur, H)
            elif L[0] == 'up':
                cur, cur_c, H = lower_up(L, cur, H)
            first_layer = False
    # ---------------- out: GN + SiLU + conv3x3 -> eps (NCHW) -----------------------------------------------------------------
    tag[0] += 1
    so = stats_slot()
    gn_stats(so, [(cur, cur_c)], H * H)
    A.need('act', npl * Bt * H * H * cur_c * H2)
    gn_apply(so, [(cur, cur_c)], H, 'out.0:g', 'out.0:b', 1e-5, 1, 'act')
    fin_c = cur_c
    emit(lambda R_: G.conv_gemm(R_('act'), Bt, R, R, fin_c, W('out.2:w'), st['out_channels'], taps=9, npass=npass, bias=W('out.2:b'),
                                edm=(0, 0, 0, st['out_channels'], io(S.DS_IO_D)))[0])
    assert stat_i[0] <= n_gn and H == R

    total = A.finalize()
    arr = (S.PlanOp * len(ops))()
    for i, (tg, builder) in enumerate(ops):
        desc = builder(A.ref)
        if isinstance(desc, S.GemmDesc) and desc.edm_out == 1 and desc.edm_x == 0:
            desc.edm_out = 2                       # plain NCHW write of eps
        arr[i].type = S.OP_TYPE_OF[type(desc)]
        arr[i].tag = tg
        setattr(arr[i].u, S.UNION_FIELD[arr[i].type], desc)
    meta = dict(B=B, Bt=Bt, nT=nT, npass=npass, ctx_tokens=ctx_tokens, n_ops=len(ops), n_gemm=sum(1 for i in range(len(ops)) if arr[i].type == S.DS_OP_GEMM))
    return Plan(arr, len(ops), total, dict(A.offsets), meta)

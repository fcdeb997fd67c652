"""Per-launch efficiency tables of one denoiser forward: GroupNorm kernels (algorithmic bytes / time) and every tcgen05 GEMM
(executed tensor flops incl. all passes and tile padding / time).   python profiles/profile_ops.py {cifar10|ffhq|imagenet64|sd15} [batch]"""
import ctypes as C
import importlib
import sys

sys.path.insert(0, '.')
import torch

importlib.import_module('diff_sampler_b200')
from diff_sampler_b200 import _cstructs as S, _lib

dev = torch.device('cuda:0')
name = sys.argv[1] if len(sys.argv) > 1 else 'cifar10'
B = int(sys.argv[2]) if len(sys.argv) > 2 else {'cifar10': 512, 'ffhq': 256, 'imagenet64': 256, 'sd15': 8}[name]


def profiled_plans():
    """-> [(plan, per-op ms list)] of one profiled forward"""
    if name == 'sd15':
        import bench

        class A:
            precision = 'fp16x3'
            num_steps = 4
            solver = 'dpm_pp'
        gen = torch.Generator(device=dev).manual_seed(1)
        net, _, kw = bench.build_sd15(A, dev, B, gen)
        x = torch.randn(B, 4, 64, 64, device=dev) * 3
        call = lambda: net(x, torch.tensor([3.0], device=dev), condition=kw['condition'], unconditional_condition=kw['unconditional_condition'])
    else:
        from diff_sampler_b200.net import B200Net
        net = B200Net.from_config(name, device=dev, seed=0, dezero=True)
        r = net.img_resolution
        x = torch.randn(B, net.img_channels, r, r, device=dev) * 10
        lab = None
        if net.label_dim:
            lab = torch.zeros(B, net.label_dim, device=dev)
            lab[torch.arange(B), torch.arange(B) % net.label_dim] = 1
        call = lambda: net(x, torch.tensor(5.0, device=dev), lab)
    call()
    call()
    plans = list(net._plans.values())
    for h, pl in plans:
        _lib.check(net.lib.ds_unet_set_profiling(h, 1), 'prof')
    call()
    out = []
    for h, pl in plans:
        buf = (C.c_float * pl.n_ops)()
        n = net.lib.ds_unet_get_profile(h, buf, pl.n_ops)
        net.lib.ds_unet_set_profiling(h, 0)
        out.append((pl, [buf[i] for i in range(n)]))
    return out, net


plans, net = profiled_plans()
names = {1: 'gemm', 2: 'gn_stats', 3: 'gn_apply', 4: 'softmax', 5: 'posemb', 6: 'linear', 7: 'prep', 8: 'chanmean', 9: 'memset', 10: 'layernorm',
         11: 'geglu', 12: 'gn_finalize', 13: 'attn'}
bytype, mem, gagg, aagg = {}, {}, {}, {}
for pl, ms_list in plans:
    for i, ms in enumerate(ms_list):
        op = pl.ops_array[i]
        a = bytype.setdefault(op.type, [0, 0.0])
        a[0] += 1; a[1] += ms
        if op.type == S.DS_OP_GN_STATS:
            d = op.u.gn_stats
            key, by = ('gn_stats', d.C0 + d.C1, d.HW, 0), d.B * d.HW * (d.C0 + d.C1) * 4
        elif op.type == S.DS_OP_GN_APPLY:
            d = op.u.gn_apply
            Cc, hw_in = d.C0 + d.C1, d.H * d.W
            hw_out = hw_in // 4 if d.resample in (1, 3) else (hw_in * 4 if d.resample == 2 else hw_in)
            by = d.B * hw_in * Cc * 4 + d.B * hw_out * Cc * 2 * d.nplanes * (2 if d.out_raw else 1) + (d.B * hw_out * Cc * 4 if d.out_raw_f32 else 0)
            key = ('gn_apply', Cc, hw_in, d.resample)
        elif op.type == S.DS_OP_SOFTMAX:
            d = op.u.softmax
            key, by = ('softmax', d.L, int(d.rows), 0), d.rows * d.L * (4 + 2 * d.nplanes)
        elif op.type == S.DS_OP_ATTN:
            d = op.u.attn
            # executed tensor flops: QK^T and PV, 3 passes each, keys padded to whole 128-blocks, head dim 64
            lk = -(-d.Lk // 128) * 128
            fl = 2.0 * 2 * 3 * d.B * d.nh * (-(-d.L // 128) * 128) * lk * 64
            k2 = ('attn', d.B * d.nh, d.L, d.Lk)
            g = aagg.setdefault(k2, [0, 0.0, 0.0])
            g[0] += 1; g[1] += fl; g[2] += ms
            continue
        elif op.type == S.DS_OP_GEMM:
            d = op.u.gemm
            kb = d.taps * d.cpb + d.nkb_aux
            fl = 2.0 * d.m_tiles * 128 * d.n_tiles * d.BN * kb * 64 * d.npass * d.num_z
            k2 = (d.m_tiles, d.n_tiles, d.BN, d.num_z, kb, d.npass, int(bool(d.st_quads)))
            g = gagg.setdefault(k2, [0, 0.0, 0.0])
            g[0] += 1; g[1] += fl; g[2] += ms
            continue
        else:
            continue
        m = mem.setdefault(key, [0, 0, 0.0])
        m[0] += 1; m[1] += by; m[2] += ms
tot = sum(v[1] for v in bytype.values())
print(f'{name} batch {B}: {tot:.3f} ms per forward over {sum(v[0] for v in bytype.values())} launches')
for t, (c, ms) in sorted(bytype.items(), key=lambda kv: -kv[1][1]):
    print(f'  {str(names.get(t, t)):12s} n={c:4d} {ms:9.3f} ms {100 * ms / tot:5.1f}%')
print()
print(f'{"kernel":9s} {"C|L":>5s} {"HW|rows":>8s} rs {"n":>3s} {"MB/launch":>10s} {"us/launch":>10s} {"GB/s":>8s}')
for (k, a, b, rs), (cnt, by, ms) in sorted(mem.items()):
    print(f'{k:9s} {a:5d} {b:8d} {rs:2d} {cnt:3d} {by / cnt / 1e6:10.1f} {ms / cnt * 1e3:10.1f} {by / (ms * 1e-3) / 1e9:8.0f}')
print()
for (k, z, L, Lk), (cnt, fl, ms) in sorted(aagg.items(), key=lambda kv: -kv[1][2]):
    print(f'attn x{cnt:<3d} heads*batch {z:5d} L {L:5d} Lk {Lk:5d}: {ms / cnt * 1e3:9.1f} us/launch  {fl / (ms * 1e-3) / 1e12:7.0f} TF/s executed  total {ms:.2f} ms')
print()
print(f'{"n":>5s} {"m_tiles":>7s} {"nt":>3s} {"BN":>4s} {"z":>5s} {"kb":>4s} {"np":>2s} {"st":>2s} {"us/launch":>10s} {"TF/s exec":>10s} {"tiles/SM":>8s} {"total ms":>9s}')
tf = tm = 0
for key, (cnt, fl, ms) in sorted(gagg.items(), key=lambda kv: -kv[1][2]):
    mt, nt, bn, z, kb, np_, st = key
    print(f'x{cnt:<4d} {mt:7d} {nt:3d} {bn:4d} {z:5d} {kb:4d} {np_:2d} {st:2d} {ms / cnt * 1e3:10.1f} {fl / (ms * 1e-3) / 1e12:10.0f} {mt * nt * z / 148:8.2f} {ms:9.2f}')
    tf += fl; tm += ms
print(f'all GEMMs: {tf / 1e12:.1f} TF executed in {tm:.2f} ms = {tf / (tm * 1e-3) / 1e12:.0f} TF/s')

"""Per-launch bytes / time / GB/s of the GroupNorm kernels of one CIFAR-10 forward (batch 512).  python profiles/profile_elementwise.py
Bytes are algorithmic: gn_stats reads the fp32 tensor once; gn_apply reads it once and writes 2 fp16 planes (+ optional raw copies)."""
import ctypes as C
import importlib
import sys

sys.path.insert(0, '.')
import torch

pkg = importlib.import_module('diff_sampler_b200')
from diff_sampler_b200 import _cstructs as S, _lib
from diff_sampler_b200.net import B200Net

dev = torch.device('cuda:0')
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
FUSE = len(sys.argv) > 2 and sys.argv[2] == 'fuse'
net = B200Net.from_config('cifar10', device=dev, seed=0, fuse_stats=FUSE)
print('fuse_stats', FUSE)
x = torch.randn(B, 3, 32, 32, device=dev) * 10
sig = torch.tensor(5.0, device=dev)
net(x, sig)
h, pl = net._plan(B, 1, 0)
lib = net.lib
for rep in range(2):
    _lib.check(lib.ds_unet_set_profiling(h, 1), 'prof')
    net(x, sig)
    buf = (C.c_float * pl.n_ops)()
    n = lib.ds_unet_get_profile(h, buf, pl.n_ops)
    lib.ds_unet_set_profiling(h, 0)
rows = []
for i in range(n):
    op = pl.ops_array[i]
    if op.type == S.DS_OP_GN_STATS:
        d = op.u.gn_stats
        by = d.B * d.HW * (d.C0 + d.C1) * 4
        rows.append(('gn_stats', d.C0 + d.C1, d.HW, 0, by, buf[i]))
    elif op.type == S.DS_OP_GN_APPLY:
        d = op.u.gn_apply
        Cc = d.C0 + d.C1
        hw_in = d.H * d.W
        hw_out = hw_in // 4 if d.resample in (1, 3) else (hw_in * 4 if d.resample == 2 else hw_in)
        by = d.B * hw_in * Cc * 4 + d.B * hw_out * Cc * 2 * d.nplanes
        if d.out_raw:
            by += d.B * hw_out * Cc * 2 * d.nplanes
        if d.out_raw_f32:
            by += d.B * hw_out * Cc * 4
        rows.append(('gn_apply', Cc, hw_in, d.resample, by, buf[i]))
agg = {}
for k, Cc, hw, rs, by, ms in rows:
    key = (k, Cc, hw, rs)
    a = agg.setdefault(key, [0, 0, 0.0])
    a[0] += 1; a[1] += by; a[2] += ms
print(f'{"kernel":9s} {"C":>5s} {"HW":>5s} rs {"n":>3s} {"MB/launch":>10s} {"us/launch":>10s} {"GB/s":>8s}')
tb = {'gn_stats': [0, 0.0], 'gn_apply': [0, 0.0]}
for (k, Cc, hw, rs), (cnt, by, ms) in sorted(agg.items()):
    print(f'{k:9s} {Cc:5d} {hw:5d} {rs:2d} {cnt:3d} {by / cnt / 1e6:10.1f} {ms / cnt * 1e3:10.1f} {by / (ms * 1e-3) / 1e9:8.0f}')
    tb[k][0] += by; tb[k][1] += ms
for k, (by, ms) in tb.items():
    if ms > 0:
        print(f'total {k}: {by / 1e9:.2f} GB in {ms:.3f} ms = {by / (ms * 1e-3) / 1e9:.0f} GB/s')

# ---- per-GEMM table: executed tensor flops (all passes, padded tiles) vs time
print()
print(f'{"gemm":5s} {"tag":>4s} {"mt":>5s} {"nt":>3s} {"BN":>4s} {"z":>4s} {"kb":>4s} {"np":>2s} {"stat":>4s} {"us":>8s} {"TF/s exec":>10s} {"tiles/SM":>8s}')
gagg = {}
for i in range(n):
    op = pl.ops_array[i]
    if op.type != S.DS_OP_GEMM:
        continue
    d = op.u.gemm
    kb = d.taps * d.cpb + d.nkb_aux
    fl = 2.0 * d.m_tiles * 128 * d.n_tiles * d.BN * kb * 64 * d.npass * d.num_z
    key = (d.m_tiles, d.n_tiles, d.BN, d.num_z, kb, d.npass, int(bool(d.st_quads)))
    a = gagg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += fl; a[2] += buf[i]
tot_fl = tot_ms = 0
for key, (cnt, fl, ms) in sorted(gagg.items(), key=lambda kv: -kv[1][2]):
    mt, nt, bn, z, kb, np_, st = key
    print(f'x{cnt:<4d} {"":>4s} {mt:5d} {nt:3d} {bn:4d} {z:4d} {kb:4d} {np_:2d} {st:4d} {ms / cnt * 1e3:8.1f} {fl / (ms * 1e-3) / 1e12:10.0f} {mt * nt * z / 148:8.2f}   total {ms:.2f} ms')
    tot_fl += fl; tot_ms += ms
print(f'all GEMMs: {tot_fl / 1e12:.1f} TF executed in {tot_ms:.2f} ms = {tot_fl / (tot_ms * 1e-3) / 1e12:.0f} TF/s')

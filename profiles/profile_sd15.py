"""Per-op timing of one SD-v1.5-sized CFG denoiser call (batch 8 -> 16 eps-net samples).  python profiles/profile_sd15.py [precision]"""
import collections
import sys

sys.path.insert(0, '.')
import torch

import bench


class A:
    precision = sys.argv[1] if len(sys.argv) > 1 else 'fp16f8'
    num_steps = 4
    solver = 'dpm_pp'


dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(1)
net, sampler, kw = bench.build_sd15(A, dev, 8, gen)
x = torch.randn(8, 4, 64, 64, device=dev) * 3
prof, per_op = net.profile_call(x, torch.tensor([3.0], device=dev), kw['condition'], kw['unconditional_condition'])
names = {1: 'gemm', 2: 'gn_stats', 3: 'gn_apply', 4: 'softmax', 5: 'posemb', 6: 'linear', 7: 'prep', 8: 'chanmean', 9: 'memset', 10: 'layernorm', 11: 'geglu',
         12: 'gn_finalize', 13: 'attn', 14: 'embed'}
tot = sum(v[1] for v in prof.values())
print('total ms per denoiser call', round(tot, 3))
for t, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f'{names.get(t, t):10s} n={c:4d} {ms:9.3f} ms {100 * ms / tot:5.1f}%')
top = sorted(per_op, key=lambda r: -r[2])[:40]
print('top ops (type, layer tag, ms):', [(names.get(t, t), tag, round(ms, 3)) for t, tag, ms in top])

# GEMMs grouped by shape: count, total ms, algorithmic TFLOP/s, BN x tiles
from diff_sampler_b200 import _cstructs as S, gemm_desc as G  # noqa: E402
groups = collections.OrderedDict()
for (h, pl) in net._plans.values():
    for i in range(pl.n_ops):
        op = pl.ops_array[i]
        if op.type != S.DS_OP_GEMM:
            continue
        d = G.describe(op.u.gemm)
        key = (d['label'], int(op.u.gemm.BN), int(op.u.gemm.m_tiles) * int(op.u.gemm.n_tiles) * max(int(op.u.gemm.num_z), 1), int(op.u.gemm.npass))
        groups.setdefault(key, []).append((pl, i, d['flops']))
ms_of = {}
idx = 0
for (h, pl) in net._plans.values():
    ms_of[id(pl)] = [r[2] for r in per_op[idx:idx + pl.n_ops]] if len(per_op) >= idx + pl.n_ops else None
    if ms_of[id(pl)] is not None and len(net._plans) == 1:
        idx += pl.n_ops
rows = []
for key, lst in groups.items():
    tot_ms = sum((ms_of[id(pl)][i] if ms_of.get(id(pl)) else 0.0) for pl, i, _ in lst)
    fl = sum(f for _, _, f in lst)
    rows.append((tot_ms, key, len(lst), fl))
print('\nGEMM shapes by total time (ms, label, BN, tiles, npass, count, algorithmic TFLOP/s):')
for tot_ms, key, cnt, fl in sorted(rows, key=lambda r: -r[0])[:45]:
    print(f'{tot_ms:8.3f} ms  {key[0]:44s} BN={key[1]:3d} tiles={key[2]:5d} npass={key[3]} n={cnt:3d}  {fl / max(tot_ms, 1e-9) / 1e9:7.1f} TF/s')

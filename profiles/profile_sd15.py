"""Per-op timing of one SD-v1.5-sized CFG denoiser call (batch 8 -> 16 eps-net samples).  python profiles/profile_sd15.py"""
import collections
import sys

sys.path.insert(0, '.')
import torch

import bench


class A:
    precision = 'fp16x3'
    num_steps = 4
    solver = 'dpm_pp'


dev = torch.device('cuda:0')
gen = torch.Generator(device=dev).manual_seed(1)
net, sampler, kw = bench.build_sd15(A, dev, 8, gen)
x = torch.randn(8, 4, 64, 64, device=dev) * 3
prof, per_op = net.profile_call(x, torch.tensor([3.0], device=dev), kw['condition'], kw['unconditional_condition'])
names = {1: 'gemm', 2: 'gn_stats', 3: 'gn_apply', 4: 'softmax', 5: 'posemb', 6: 'linear', 7: 'prep', 8: 'chanmean', 9: 'memset', 10: 'layernorm', 11: 'geglu'}
tot = sum(v[1] for v in prof.values())
print('total ms per denoiser call', round(tot, 3))
for t, (c, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]):
    print(f'{names.get(t, t):10s} n={c:4d} {ms:9.3f} ms {100 * ms / tot:5.1f}%')
top = sorted(per_op, key=lambda r: -r[2])[:25]
print('top ops (type, layer tag, ms):', [(names.get(t, t), tag, round(ms, 3)) for t, tag, ms in top])

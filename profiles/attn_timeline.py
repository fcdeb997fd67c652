"""Timeline of one CTA of the fused attention kernel (ds_debug_attn_trace): which role waits for what, in SM cycles.
    python profiles/attn_timeline.py [L] [B] [nh] > gpurun_out/<run>/attn_timeline.txt
Roles: TMA (K / V block issued), MMA (qk: wait K | wait S free | issued;  pv: wait V | wait P | issued), G<g> softmax group g
(wait S | got S | max done | ready to write P | P written | arrived)."""
import ctypes as C
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_sampler_b200 import _cstructs as S, _lib  # noqa: E402
from diff_sampler_b200.gemm_desc import split_planes  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nh = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = torch.device('cuda:0')
lib = _lib.load()
torch.manual_seed(0)
hp = nh * 64
qk = split_planes(torch.randn(B, L, 2 * hp, device=dev))
vt = split_planes(torch.randn(B, hp, L, device=dev))
out = torch.zeros(2, B, L, hp, dtype=torch.float16, device=dev)
d = S.AttnDesc(q=qk.data_ptr(), k=qk.data_ptr(), vt=vt.data_ptr(), out=out.data_ptr(), B=B, nh=nh, L=L, Lk=L, q_pitch=2 * hp, q_c0=0, k_pitch=2 * hp,
               k_c0=hp, vt_pitch=L, o_pitch=hp, nplanes=2, scale=0.125)
for _ in range(3):
    _lib.op_launch(d)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    _lib.op_launch(d)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
blocks = B * nh * ((L + 127) // 128) * ((L + 63) // 64)
print(f'# attention B={B} heads={nh} L={L}: {ms * 1e3:.1f} us per launch; {blocks} key blocks -> {ms * 1e-3 / blocks * 148 * 1.8e9:.0f} cycles per block per SM at 1.8 GHz')
cap = 1 << 15
buf = torch.zeros(cap, dtype=torch.int64, device=dev)
lib.ds_debug_attn_trace(buf.data_ptr(), cap)
_lib.op_launch(d)
torch.cuda.synchronize()
lib.ds_debug_attn_trace(None, 0)
host = buf.cpu().tolist()
per = cap // 8
ev = []
for who in range(8):                      # every role has its own region: [0] = events written + 1
    cnt = host[who * per]
    ev += host[who * per + 1: who * per + max(cnt, 1)]
n = len(ev)
rows = []
for e in ev:
    tag, clk = (e >> 40) & 0xFFFFFF, e & 0xFFFFFFFFFF
    rows.append((clk, (tag >> 20) & 15, (tag >> 16) & 15, (tag >> 8) & 255, tag & 255))
rows.sort()
t0 = rows[0][0]
NAMES = {0: {0: 'K issued', 1: 'V issued'},
         1: {0: 'qk wait K', 1: 'qk got K', 2: 'qk S free -> issue', 4: 'pv wait V', 5: 'pv got V', 6: 'pv got P', 7: 'pv issued'}}
SM = {0: 'wait S', 1: 'got S', 2: 'max done', 3: 'P buffer free', 4: 'P written', 5: 'fenced'}
print(f'# {n} events; cycles relative to the first; tile, role, event, key block')
for clk, it, who, e, j in rows:
    role = 'TMA' if who == 0 else ('MMA' if who == 1 else f'G{who - 2}')
    name = NAMES.get(who, SM).get(e, str(e))
    print(f'{clk - t0:9d}  tile{it} {role:4s} {name:22s} j={j}')

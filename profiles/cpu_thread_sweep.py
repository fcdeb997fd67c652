import sys, time, os
sys.path.insert(0, '.')
import torch
from oracle import edm_oracle as O
P, S = O.make_net('cifar10', seed=0, dezero=True)
net = O.OracleNet(P, S)
x = O.stacked_randn(range(8), (3, 32, 32))
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for th in (8, 16, 32, 64, 128):
    torch.set_num_threads(th)
    net(x, torch.tensor(2.0))
    t0 = time.perf_counter(); net(x, torch.tensor(2.0)); dt = time.perf_counter() - t0
    print('threads', th, 'fwd s', round(dt, 3), flush=True)

"""Timeline of the conv GEMM's shared-memory ring on CTA 0 (ds_debug_gemm_trace): per ring stage, the SM clock at which the TMA producer got
its empty slot (t_p) and at which the MMA warp saw the slot full (t_f).  Prints issue intervals, load latencies (t_f - t_p) and how far
the producer runs ahead, for the CIFAR-10 256->256 32x32 shape (f8 mode), single-CTA and pair kernels, optionally under DSB_GEMM_DIAG modes.

    python profiles/gemm_timeline.py [--diag 0,5] > gpurun_out/<run>/gemm_timeline.txt
"""
import argparse
import os
import sys

import torch

os.environ['DSB_GEMM_2CTA'] = '0'
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_sampler_b200 import _lib, gemm_desc as G  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--diag', default='0,5')
    ap.add_argument('--cap', type=int, default=1024)
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    Bn, H, W, Cin, Cout = 512, 32, 32, 256, 256
    x = torch.randn(Bn, H, W, Cin, device=dev)
    w = torch.randn(Cout, Cin, 3, 3) / (3 * Cin ** 0.5)
    out = torch.empty(Bn * H * W, Cout, device=dev)
    blob, shift = G.pack_conv_weight_f8(w)
    xa = G.act_planes_f8(x)
    wp = blob.to(dev)
    cap = args.cap
    buf = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
    for pair in (False, True):
        for diag in args.diag.split(','):
            os.environ['DSB_GEMM_DIAG'] = diag
            lib.ds_debug_gemm_trace(None, 0)
            d, _ = G.conv_gemm(xa.data_ptr(), Bn, H, W, Cin, wp.data_ptr(), Cout, taps=9, npass=3, out_f32=out.data_ptr(), bn=256, pair=pair,
                               f8=True, acc_scale=2.0 ** -shift)
            for _ in range(3):
                _lib.op_launch(d)
            torch.cuda.synchronize()
            buf.zero_()
            lib.ds_debug_gemm_trace(buf.data_ptr(), 2 * cap)
            _lib.op_launch(d)
            torch.cuda.synchronize()
            lib.ds_debug_gemm_trace(None, 0)
            t = buf.cpu().tolist()
            tp, tf = t[:cap], t[cap:]
            n = min(sum(1 for v in tp if v), sum(1 for v in tf if v))
            t0 = tp[0]
            tp = [v - t0 for v in tp[:n]]
            tf = [v - t0 for v in tf[:n]]
            print(f'== cifar 32^2 256->256 f8 BN=256 pair={int(pair)} diag={diag}: {n} stages traced (72 per tile)')
            lat = sorted(tf[i] - tp[i] for i in range(144, n))
            iv = sorted(tf[i + 1] - tf[i] for i in range(144, n - 1))
            ahead = []
            for i in range(144, n):
                j = i
                while j + 1 < n and tp[j + 1] < tf[i]:
                    j += 1
                ahead.append(j - i)
            ahead.sort()
            q = lambda a, f: a[int(f * (len(a) - 1))]
            print(f'   steady state (stage 144..): full-to-full interval median {q(iv, .5)} (p10 {q(iv, .1)}, p90 {q(iv, .9)}) cycles; '
                  f'slot granted -> slot full median {q(lat, .5)} (p10 {q(lat, .1)}, p90 {q(lat, .9)}); producer ahead of the MMA warp by median {q(ahead, .5)} stages '
                  f'(p10 {q(ahead, .1)}, p90 {q(ahead, .9)})')
            print('   stage: t_producer  t_full   (first 40 stages, then stages 144..183)')
            for i in list(range(0, 40)) + list(range(144, min(184, n))):
                print(f'   {i:4d}: {tp[i]:9d} {tf[i]:9d}   lat {tf[i] - tp[i]:6d}   d_full {tf[i] - tf[i - 1] if i else 0:6d}')
    os.environ['DSB_GEMM_DIAG'] = '0'


if __name__ == '__main__':
    main()

"""Microbenchmark of the tcgen05 conv kernel over N-tile choices for the convolution shapes of the ImageNet-64 ADM net (C = 192 .. 768) and
of the small-M SD-v1.5 levels: time per launch (CUDA events, back-to-back launches long enough for the power cap to settle) and
executed TFLOP/s, for fp16x3 and the f8 mode, single-CTA and CTA-pair kernels.  Decides gemm_desc.pick_bn / fill_bn.

    python profiles/bench_gemm_tiles.py > gpurun_out/<run>/gemm_tiles.txt
"""
import sys
import os
import time

import torch

os.environ['DSB_GEMM_2CTA'] = '0'          # the pair kernel only where this script asks for it (conv_gemm(pair=True))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from diff_sampler_b200 import _lib, gemm_desc as G  # noqa: E402

dev = torch.device('cuda:0')
SHAPES = [  # (label, Bn, H, W, Cin, Cout)
    ('adm 64^2 192->192', 256, 64, 64, 192, 192), ('adm 32^2 384->384', 256, 32, 32, 384, 384), ('adm 16^2 576->576', 256, 16, 16, 576, 576),
    ('adm 8^2 768->768', 256, 8, 8, 768, 768), ('sd 8^2 1280->1280 x16', 16, 8, 8, 1280, 1280), ('sd 16^2 1280->1280 x16', 16, 16, 16, 1280, 1280),
    ('sd 32^2 640->640 x16', 16, 32, 32, 640, 640), ('sd 64^2 320->320 x16', 16, 64, 64, 320, 320), ('cifar 32^2 256->256', 512, 32, 32, 256, 256),
    # short-K 1x1 convolutions / linears of the SD transformer blocks (7th field: taps)
    ('sd1x1 64^2 320->2560 x16', 16, 64, 64, 320, 2560, 1), ('sd1x1 64^2 320->320 x16', 16, 64, 64, 320, 320, 1),
    ('sd1x1 64^2 512->320 x16', 16, 64, 64, 512, 320, 1), ('sd1x1 32^2 640->5120 x16', 16, 32, 32, 640, 5120, 1),
]


_nvml = None
LAST_MHZ = 0


def sm_mhz():
    global _nvml
    try:
        import pynvml
        if _nvml is None:
            pynvml.nvmlInit()
            _nvml = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())
        return pynvml.nvmlDeviceGetClockInfo(_nvml, pynvml.NVML_CLOCK_SM)
    except Exception:
        return 0


def timed(d, n):
    global LAST_MHZ
    _lib.op_launch(d)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        _lib.op_launch(d)
    e1.record()
    clocks = []
    while not e1.query():                      # SM clock while the launches drain (power cap: the clock under THIS kernel's load)
        clocks.append(sm_mhz())
        time.sleep(0.01)
    torch.cuda.synchronize()
    LAST_MHZ = sorted(clocks)[len(clocks) // 2] if clocks else 0
    return e0.elapsed_time(e1) / n


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default='', help='substring of the shape label (for ncu captures of one shape)')
    ap.add_argument('--bn', type=int, default=0, help='only this N tile')
    ap.add_argument('--mode', default='', help='x3 | f8 (default both)')
    ap.add_argument('--reps', type=int, default=0, help='fixed number of timed launches (default: ~0.25 s worth)')
    ap.add_argument('--diag', default='0', help='comma list of DSB_GEMM_DIAG modes (1 = no MMA, 2 = no TMA loads, 4 = no epilogue; sums allowed): '
                                                 'which of operand feed / MMA issue / epilogue bounds the kernel')
    args = ap.parse_args()
    _lib.load()
    torch.manual_seed(0)
    for label, Bn, H, W, Cin, Cout, *rest in SHAPES:
        if args.only and args.only not in label:
            continue
        taps = rest[0] if rest else 9
        ks = 3 if taps == 9 else 1
        x = torch.randn(Bn, H, W, Cin, device=dev)
        w = torch.randn(Cout, Cin, ks, ks) / (ks * Cin ** 0.5)
        out = torch.empty(Bn * H * W, Cout, device=dev)
        flops = 2.0 * Bn * H * W * Cout * Cin * taps
        m_tiles = -(-(Bn * H * W) // 128)
        auto_bn, _ = G.fill_bn(Cout, m_tiles)
        cands = sorted({c for c in (64, 80, 96, 128, 144, 160, 192, 256, auto_bn) if c <= 256 and (Cout % c == 0 or c == auto_bn or Cout > 256)})
        if args.bn:
            cands = [args.bn]
        for f8 in (False, True):
            if args.mode and args.mode != ('f8' if f8 else 'x3'):
                continue
            if f8:
                blob, shift = G.pack_conv_weight_f8(w)
                xa = G.act_planes_f8(x)
                kw = dict(f8=True, acc_scale=2.0 ** -shift)
            else:
                blob = G.pack_conv_weight(w)
                xa = G.split_planes(x)
                kw = {}
            wp = blob.to(dev)
            for bn in cands:
                for pair in (False, True):
                    if pair and (bn % 32 or m_tiles < 2):
                        continue
                    for diag in args.diag.split(','):
                        os.environ['DSB_GEMM_DIAG'] = diag
                        try:
                            d, _ = G.conv_gemm(xa.data_ptr(), Bn, H, W, Cin, wp.data_ptr(), Cout, taps=taps, npass=3, out_f32=out.data_ptr(), bn=bn, pair=pair, **kw)
                            n = max(20, int(0.25 / max(flops * (2 if f8 else 3) / 1.2e15, 1e-5)))     # ~0.25 s of launches
                            ms = timed(d, args.reps if args.reps else min(n, 2000))
                        except Exception as e:
                            print(f'{label:24s} {"f8 " if f8 else "x3 "} BN={bn:3d} pair={int(pair)}  FAILED {e!r}'[:160])
                            continue
                        tiles = m_tiles * -(-Cout // bn)
                        mark = ' <- fill_bn' if bn == auto_bn else ''
                        if diag != '0':
                            mark += f'  [diag {diag}: ' + '+'.join(n_ for b_, n_ in ((1, 'no MMA'), (2, 'no TMA'), (4, 'no epilogue')) if int(diag) & b_) + ']'
                        print(f'{label:24s} {"f8 " if f8 else "x3 "} BN={bn:3d} pair={int(pair)} tiles={tiles:6d} ({tiles / 148:6.2f}/SM)  {ms * 1e3:9.1f} us  '
                              f'{flops / ms / 1e9:7.1f} TF/s algorithmic  {flops * (2 if f8 else 3) / ms / 1e9:7.1f} executed  {LAST_MHZ:4d} MHz{mark}', flush=True)
                    os.environ['DSB_GEMM_DIAG'] = '0'
        del x, out
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()

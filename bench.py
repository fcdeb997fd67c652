#!/usr/bin/env python
"""bench.py — images/sec at fixed NFE for the BASELINE headline configuration.

    python bench.py --gpus N --steps K --warmup W            # native arm (one process per GPU under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path, timed on the host cores (rank 0 only)

Workload (BASELINE.json configs[1]): EDM CIFAR-10 32x32 DDPM++ U-Net (random init, de-zeroed so |F_x| = O(1)), Heun sampler,
num_steps=10 => NFE=18, batch 512 per GPU, synthetic Gaussian latents.  One "step" = one full sampling pass over one batch.
`value` = images/sec with latents resident in HBM; `e2e` = the same through the public API with pinned-host latents copied in
and finished images copied back every step.  Scaling is weak: every rank samples its own 512-image batch, no collective on
the sampling path; one NCCL all_gather of the uint8 images after the timed region (what FID consumes).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GFLOP_PER_IMG_NFE = {'cifar10': 42.38, 'ffhq': 83.73, 'imagenet64': 219.33}      # BASELINE.md section 2 (2 FLOP per MAC)
SOLVER_NFE = {'heun': lambda n: 2 * (n - 1), 'euler': lambda n: n - 1, 'ipndm': lambda n: n - 1, 'dpm_pp': lambda n: n - 1}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='native', choices=['native', 'reference'])
    ap.add_argument('--net', default='cifar10')
    ap.add_argument('--solver', default='heun')
    ap.add_argument('--num_steps', type=int, default=10)
    ap.add_argument('--batch', type=int, default=512, help='images per GPU per step')
    ap.add_argument('--precision', default='fp16x3', choices=['fp16x3', 'fp16'])
    ap.add_argument('--cpu_batch', type=int, default=8, help='batch of the bounded CPU-baseline sample')
    ap.add_argument('--no_cpu_baseline', action='store_true')
    ap.add_argument('--no_extras', action='store_true', help='skip the roofline / e2e / fp16 legs (timing of the main leg is unchanged)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm_gbs=d['hbm_gbs'], tflops_burst=d['bf16_tflops'], tflops_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200', '-i', str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['nvidia-smi unavailable'])
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return dict(sm_mhz=med, sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm))


# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_leg(args, steps, warmup):
    """The reference's CPU path on the host cores: the oracle port of solvers.<solver>_sampler on EDMPrecond (oracle/ is pinned
    bit-exact to /root/reference by tests/golden; /root/reference itself does not exist on the GPU box).  Bounded sample:
    the same net / solver / NFE on a batch of `cpu_batch` images."""
    import torch
    from oracle import edm_oracle as O
    from oracle import solvers_oracle as SO
    # thread count: measured on the B200 host (profiles/cpu_thread_sweep.py, 128 logical CPUs): one CIFAR-net forward at batch 8
    # takes 0.21 s with 16 threads, 0.27 s with 32, 0.59 s with 64 and 4.6 s with 128 (oversubscription), so the baseline uses
    # the fastest setting rather than every logical CPU.
    cores = min(os.cpu_count() or 1, int(os.environ.get('DSB_CPU_THREADS', '16')))
    torch.set_num_threads(cores)
    P, S = O.make_net(args.net, seed=0, dezero=True)
    net = O.OracleNet(P, S)
    lat = O.stacked_randn(range(args.cpu_batch), (S['img_channels'], S['img_resolution'], S['img_resolution']))
    lab = None
    if S['label_dim']:
        lab = torch.eye(S['label_dim'])[torch.arange(args.cpu_batch) % S['label_dim']]
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        SO.sample(net, lat, args.solver, class_labels=lab, num_steps=args.num_steps)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = sum(times)
    return dict(value=args.cpu_batch * len(times) / tot, seconds_per_step=tot / len(times), cores=cores,
                sample=f'{args.net} {args.solver} num_steps={args.num_steps} batch {args.cpu_batch} (same net/solver/NFE, bounded batch), '
                       f'{len(times)} timed + {warmup} warm-up passes, torch CPU fp32 {torch.__version__}, {cores} threads')


def main():
    args = parse()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    nfe = SOLVER_NFE[args.solver](args.num_steps)
    config = dict(workload=f'EDM {args.net} U-Net, {args.solver} num_steps={args.num_steps} (NFE={nfe}), batch {args.batch}/GPU',
                  net=args.net, solver=args.solver, nfe=nfe, batch_per_gpu=args.batch, global_batch=args.batch * max(world, 1),
                  weights='random init (reference constructors, seed 0), init_zero layers de-zeroed', parallelism=f'dp{world}',
                  l2='per-forward activation working set (GBs) >> 126 MB L2')
    metric = 'images/sec at fixed NFE'

    if args.impl == 'reference':
        if rank != 0:
            return
        cb = cpu_reference_leg(args, max(1, args.steps), min(args.warmup, 1))
        line = dict(metric=metric, value=cb['value'], unit='images/s', n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                    ms_per_step=cb['seconds_per_step'] * 1e3, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                    data='synthetic', impl='reference', config=config,
                    cpu_baseline=dict(value=cb['value'], unit='images/s', cores=cb['cores'], kind='port', sample=cb['sample']),
                    e2e=dict(value=cb['value'], unit='images/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from diff_sampler_b200 import solver_utils, solvers
    from diff_sampler_b200.net import B200Net
    assert torch.cuda.is_available(), 'the native arm needs a CUDA device (no CPU fallback)'
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # NCCL prints its version banner on stdout at communicator creation; stdout must carry exactly one JSON line.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    net = B200Net.from_config(args.net, seed=0, dezero=True, precision=args.precision, device=dev)
    sampler = getattr(solvers, args.solver + '_sampler')
    B = args.batch
    shape = (B, net.img_channels, net.img_resolution, net.img_resolution)
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    latents = torch.randn(shape, generator=gen, device=dev)
    labels = None
    if net.label_dim:
        labels = torch.eye(net.label_dim, device=dev)[torch.randint(net.label_dim, (B,), generator=gen, device=dev)]
    kw = dict(class_labels=labels, num_steps=args.num_steps, sigma_min=0.002, sigma_max=80, schedule_type='polynomial', schedule_rho=7)

    def run_step():
        return sampler(net, latents, **kw)

    for _ in range(args.warmup):
        run_step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = net.total_launches + solver_utils.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        images = run_step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = net.total_launches + solver_utils.LAUNCHES[0] - l0
    clk = clocks.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * args.steps / (ms / 1e3)

    # ---- end-to-end through the public API with host buffers -------------------------------------------------------
    e2e = None
    if not args.no_extras:
        host_in = torch.randn(shape).pin_memory()
        host_out = torch.empty(shape).pin_memory()
        dev_in = torch.empty(shape, device=dev)

        def e2e_step():
            dev_in.copy_(host_in, non_blocking=True)
            out = sampler(net, dev_in, **kw)
            host_out.copy_(out, non_blocking=True)
        for _ in range(2):
            e2e_step()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(args.steps):
            e2e_step()
        a1.record()
        barrier()
        t2 = torch.tensor([a0.elapsed_time(a1)], device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        nbytes = host_in.numel() * 4
        e2e = dict(value=world * B * args.steps / (t2.item() / 1e3), unit='images/s', h2d_bytes_per_step=nbytes, d2h_bytes_per_step=nbytes)

    # ---- finished samples: one NCCL all_gather of the uint8 images (outside the timed region) -------------------------
    gathered = None
    if world > 1:
        u8 = (images * 127.5 + 128).clip(0, 255).to(torch.uint8)
        allimg = [torch.empty_like(u8) for _ in range(world)]
        dist.all_gather(allimg, u8)
        gathered = sum(x.numel() for x in allimg)

    if rank != 0:
        if world > 1:
            dist.barrier()                       # stay alive until rank 0 has finished its single-GPU diagnostic legs
            dist.destroy_process_group()
        return

    pk = peaks()
    line = dict(metric=metric, value=value, unit='images/s', n_gpus=world, steps=args.steps, warmup=args.warmup,
                ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='fp16 operands, fp32 accumulate' + (' (split-precision: 3 tcgen05 MMAs per product)' if args.precision == 'fp16x3' else ''),
                data='synthetic', config=config, gpu_launches=launches, clocks=clk, precision=args.precision)
    if e2e:
        line['e2e'] = e2e
    if gathered:
        line['allgather_bytes'] = gathered

    try:
        if not args.no_extras:
            extras(args, line, net, sampler, kw, latents, labels, images, B, dev, pk)
    except Exception as e:                       # the main measurement above is already complete; report instead of dying
        line['extras_error'] = repr(e)

    if not args.no_cpu_baseline and world == 1:
        cb = cpu_reference_leg(args, 1, 1)
        line['cpu_baseline'] = dict(value=cb['value'], unit='images/s', cores=cb['cores'], kind='port', sample=cb['sample'])
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras(args, line, net, sampler, kw, latents, labels, images, B, dev, pk):
    import torch
    from diff_sampler_b200 import solver_utils
    from diff_sampler_b200.net import B200Net
    if True:
        # ---- roofline of the dominant kernel (tcgen05 GEMM/conv), measured live with CUDA events on the launch stream ---
        x = latents * 2.0
        sig = torch.tensor(2.0, device=dev)
        prof = net.profile_forward(x, sig, labels)
        from diff_sampler_b200 import _cstructs as S
        gemm_n, gemm_ms = prof.get(S.DS_OP_GEMM, (0, 0.0))
        fwd_ms = sum(v[1] for v in prof.values())
        flops = GFLOP_PER_IMG_NFE.get(args.net, 0.0) * 1e9 * B
        achieved = flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        peak = pk['tflops_sustained']
        line['roofline'] = dict(bound='tensor', achieved=achieved, peak=peak, unit='TFLOP/s', frac=(achieved / peak) if achieved else None,
                                traffic=None, kernel='gemm_tc_kernel (all conv/attention contractions of one forward)',
                                algorithmic_flops_per_forward=flops, launches_per_forward=gemm_n, gemm_ms_per_forward=gemm_ms,
                                all_ops_ms_per_forward=fwd_ms, gemm_share_of_forward=gemm_ms / fwd_ms if fwd_ms else None,
                                executed_mma_flops_factor=3 if args.precision == 'fp16x3' else 1, peak_source=pk['source'] + ', sustained bf16 GEMM')
        line['forward_breakdown_ms'] = {str(k): round(v[1], 4) for k, v in sorted(prof.items())}
        # ---- the fused solver-update kernel against the HBM roofline (HBM-resident size: 3 x 1 GiB streams) -------------
        n = 256 * 1024 * 1024
        a, b_, c = (torch.empty(n, device=dev).normal_() for _ in range(3))
        a, b_, c = a.view(1024, -1), b_.view(1024, -1), c.view(1024, -1)
        for _ in range(3):
            solver_utils.solver_update(c, a, [1.0, 0.3], mode=S.DS_M_EPS, D=b_, t=2.0)
        torch.cuda.synchronize()
        u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        u0.record()
        for _ in range(10):
            solver_update_out = solver_utils.solver_update(c, a, [1.0, 0.3], mode=S.DS_M_EPS, D=b_, t=2.0)
        u1.record()
        torch.cuda.synchronize()
        gbs = 3 * n * 4 * 10 / (u0.elapsed_time(u1) / 1e3) / 1e9
        line['roofline_update'] = dict(bound='hbm', achieved=gbs, peak=pk['hbm_gbs'], unit='GB/s', frac=gbs / pk['hbm_gbs'],
                                       kernel='update_kernel<0, EPS> (Euler step: read x, D; write x+)', bytes_per_launch=3 * n * 4,
                                       peak_source=pk['source'])
        del a, b_, c
        # ---- single-pass fp16 (reported, not the headline: misses the 1e-3 gate on O(1) random nets) -------------------
        if args.precision == 'fp16x3':
            net1 = B200Net.from_config(args.net, seed=0, dezero=True, precision='fp16', device=dev)
            for _ in range(2):
                sampler(net1, latents, **kw)
            torch.cuda.synchronize()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(args.steps):
                img1 = sampler(net1, latents, **kw)
            f1.record()
            torch.cuda.synchronize()
            line['fp16_single_pass'] = dict(value=B * args.steps / (f0.elapsed_time(f1) / 1e3), unit='images/s (1 GPU)',
                                            max_abs_vs_fp16x3=(img1 - images).abs().max().item(),
                                            note='single tcgen05 pass per product; not the headline because it does not hold 1e-3 on the de-zeroed weight set')
            del net1



if __name__ == '__main__':
    main()

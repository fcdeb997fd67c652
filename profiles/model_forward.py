"""Where one denoiser forward spends its time versus what the hardware allows (no GPU needed: the plan is compiled on the host).

    python profiles/model_forward.py [net=cifar10] [batch=512] [precision=fp16f8] [bench.json]

For every op of the plan: algorithmic FLOPs / bytes (gemm_desc.describe for GEMMs, tensor sizes for the elementwise ops), converted to a
floor time with the MEASURED ceilings of this pool (MEASURED_PEAKS.json: sustained bf16 GEMM rate, copy bandwidth):
  GEMM        executed MMA units / sustained rate, units = 3 (fp16x3) or 2 (f8 GEMMs: e4m3 MMAs at twice the rate) per product
  GN apply    (4 B read + 4 B write) per element / copy bandwidth;   GN stats 4 B read;   softmax 8 B per score
If a bench line (profiles/r01*/bench_*.json) is given, its forward_breakdown_ms is printed beside the floors.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from diff_sampler_b200 import _cstructs as S
    from diff_sampler_b200 import edm_nets, gemm_desc as G, plan as planner
    net = sys.argv[1] if len(sys.argv) > 1 else 'cifar10'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    prec = sys.argv[3] if len(sys.argv) > 3 else 'fp16f8'
    bench = sys.argv[4] if len(sys.argv) > 4 else None
    pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')) else {}
    tf, bw = pk.get('bf16_tflops_sustained', 1400.0) * 1e12, pk.get('hbm_gbs', 6650.0) * 1e9
    params, cfg = edm_nets.init_params(net, seed=0)
    spec = edm_nets.spec_from_params(params, cfg['img_resolution'], cfg['img_channels'], cfg.get('label_dim', 0))
    f8 = prec == 'fp16f8'
    wb, info = planner.pack_weights(spec, params, f8=f8)
    pl = planner.compile_plan(spec, wb, info, B, 1, B if spec.label_dim else 0, npass=1 if prec == 'fp16' else 3, f8=f8)
    rows = {}

    def add(kind, n, flops, bytes_, floor):
        r = rows.setdefault(kind, [0, 0.0, 0.0, 0.0])
        r[0] += n; r[1] += flops; r[2] += bytes_; r[3] += floor
    for i in range(pl.n_ops):
        op = pl.ops_array[i]
        if op.type == S.DS_OP_GEMM:
            d = G.describe(op.u.gemm)
            units = 1 if prec == 'fp16' else (2 if d['f8'] else 3)
            add('gemm (f8)' if d['f8'] else 'gemm', 1, d['flops'], d['bytes'], max(units * d['flops'] / tf, d['bytes'] / bw))
        elif op.type == S.DS_OP_GN_APPLY:
            g = op.u.gn_apply
            C, rs = g.C0 + g.C1, g.resample
            n_in = g.B * g.H * g.W * C
            n_out = n_in // 4 if rs == 1 else (n_in * 4 if rs == 2 else n_in)
            nbytes = 4 * n_in + 4 * n_out * ((1 if g.out_act else 0) + (1 if g.out_raw else 0)) + (4 * n_out if g.out_raw_f32 else 0)
            add('gn_apply', 1, 0, nbytes, nbytes / bw)
        elif op.type == S.DS_OP_GN_STATS:
            g = op.u.gn_stats
            nbytes = 4 * g.B * g.HW * (g.C0 + g.C1)
            add('gn_stats', 1, 0, nbytes, nbytes / bw)
        elif op.type == S.DS_OP_GN_FINALIZE:
            g = op.u.gn_finalize
            nbytes = 4 * g.B * g.slabs_per_sample * (g.C0 + g.C1) // 2
            add('gn_finalize', 1, 0, nbytes, nbytes / bw)
        elif op.type == S.DS_OP_SOFTMAX:
            g = op.u.softmax
            add('softmax', 1, 0, 8 * g.rows * g.L, 8 * g.rows * g.L / bw)
        elif op.type == S.DS_OP_ATTN:
            g = op.u.attn
            fl = 4.0 * g.B * g.nh * g.L * g.Lk * 64
            add('attn (fused)', 1, fl, 0, 3 * fl / tf)
        else:
            add('other', 1, 0, 0, 0.0)
    measured = {}
    if bench:
        line = json.loads(open(bench).read().strip().splitlines()[-1])
        fb = line.get('forward_breakdown_ms', {})
        measured = {'gemm': fb.get('1', 0.0), 'gn_apply': fb.get('3', 0.0), 'gn_stats': fb.get('2', 0.0), 'gn_finalize': fb.get('12', 0.0),
                    'softmax': fb.get('4', 0.0), 'attn (fused)': fb.get('13', 0.0)}
    print(f'# {net} batch {B} {prec}: one forward, floors at {tf / 1e12:.0f} TFLOP/s (sustained bf16 GEMM) and {bw / 1e9:.0f} GB/s (copy)')
    print(f'{"op":14s} {"n":>4s} {"TFLOP":>9s} {"GB":>8s} {"floor ms":>9s} {"measured ms":>12s}')
    tot = 0.0
    gem = sum(v[3] for k, v in rows.items() if k.startswith('gemm')) * 1e3
    for k, (n, fl, by, t) in sorted(rows.items(), key=lambda kv: -kv[1][3]):
        m = measured.get(k)
        if k == 'gemm (f8)' and 'gemm' in measured:
            m = None
        tot += t
        print(f'{k:14s} {n:4d} {fl / 1e12:9.2f} {by / 1e9:8.2f} {t * 1e3:9.2f} {"" if m is None else format(m, "12.2f")}')
    print(f'{"total":14s} {"":4s} {"":9s} {"":8s} {tot * 1e3:9.2f} {sum(measured.values()) if measured else 0:12.2f}   (all GEMMs: floor {gem:.2f} ms, measured {measured.get("gemm", 0):.2f} ms)')


if __name__ == '__main__':
    main()

"""solver_utils — drop-in for the reference module of the same name
(diff-solvers-main/solver_utils.py, gits-main/solver_utils.py, amed-solver-main/solver_utils.py).

Same public names and argument meaning.  Schedules and coefficient tables are tiny host-side computations
(kept in torch on the CPU so the fp32 values are device-independent); everything that touches an image-sized
tensor goes through the C ABI kernels (ds_solver_update / ds_dyn_threshold).
"""
import ctypes as C
import math

import numpy as np
import torch

from . import _cstructs as S
from . import _lib

__all__ = ['get_schedule', 'expand_dims', 'dynamic_thresholding_fn', 'dpm_pp_update', 'dpm_solver_first_update',
           'multistep_dpm_solver_second_update', 'multistep_dpm_solver_third_update', 'unipc_update', 'edm2t', 'cal_poly',
           't2alpha_fn', 'cal_intergrand', 'get_deis_coeff_list', 'solver_update', 'dyn_threshold']


# ---------------------------------------------------------------------------------------------------------------------
# time schedules (reference: solver_utils.py:6-52; dp_list variant gits-main/solver_utils.py:6,52-53)

def get_schedule(num_steps, sigma_min, sigma_max, device=None, schedule_type='polynomial', schedule_rho=7, net=None, dp_list=None):
    """Returns the fp32 sigma grid [num_steps] on `device`.  Computed on the host so that every rank / device sees
    bit-identical values; `dp_list` (GITS) gathers an integer index list from the fine teacher grid."""
    if schedule_type == 'polynomial':
        idx = torch.arange(num_steps)
        hi, lo = sigma_max ** (1 / schedule_rho), sigma_min ** (1 / schedule_rho)
        t_steps = (hi + idx / (num_steps - 1) * (lo - hi)) ** schedule_rho
    elif schedule_type == 'logsnr':
        lmax = (-1 * torch.log(torch.tensor(sigma_min))).item()
        lmin = (-1 * torch.log(torch.tensor(sigma_max))).item()
        t_steps = (-torch.linspace(lmin, lmax, steps=num_steps)).exp()
    elif schedule_type == 'time_uniform':
        eps_s = 1e-3
        ln_min, ln_max = np.log(torch.tensor(sigma_min) ** 2 + 1), np.log(torch.tensor(sigma_max) ** 2 + 1)
        beta_d = 2 * (ln_min / eps_s - ln_max) / (eps_s - 1)
        beta_min = ln_max - 0.5 * beta_d
        idx = torch.arange(num_steps)
        tau = (1 + idx / (num_steps - 1) * (eps_s ** (1 / schedule_rho) - 1)) ** schedule_rho
        t_steps = (np.e ** (0.5 * beta_d * (tau ** 2) + beta_min * tau) - 1) ** 0.5
    elif schedule_type == 'discrete':
        assert net is not None
        t_lo = net.sigma_inv(torch.tensor(sigma_min, device=device))
        t_hi = net.sigma_inv(torch.tensor(sigma_max, device=device))
        idx = torch.arange(num_steps, device=device)
        t_steps = net.sigma((t_hi + idx / (num_steps - 1) * (t_lo ** (1 / schedule_rho) - t_hi)) ** schedule_rho)
    else:
        raise ValueError("Got wrong schedule type {}".format(schedule_type))
    if dp_list is not None:
        return t_steps[dp_list].to(device)
    return t_steps.to(device)


def expand_dims(v, dims):
    return v[(...,) + (None,) * (dims - 1)]


# ---------------------------------------------------------------------------------------------------------------------
# kernel front-ends

LAUNCHES = [0]      # kernels launched through this module (bench.py's gpu_launches)


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _chk_dev(*ts):
    for t in ts:
        if t is not None and (t.device.type != 'cuda' or t.dtype != torch.float32 or not t.is_contiguous()):
            raise _lib.DsError('solver kernels need contiguous fp32 CUDA tensors (no CPU fallback)')


def solver_update(out_x, xb, coef, *, mode=S.DS_M_NONE, D=None, xs=None, hist=(), out_m=None, thr=None, t=1.0, t_dev=None,
                  coef_dev=None, out_u8=None):
    """out_x = coef[0]*xb + coef[1]*m0 + sum_k coef[2+k]*hist[k], with m0 derived per `mode` (include/diffsampler_b200.h).
    One kernel launch, one pass over HBM.  out_u8 ([B, H, W, C] uint8): also write the finished images as sample.py:311 does
    ((x * 127.5 + 128).clip(0, 255).uint8, NHWC) in the same pass."""
    lib = _lib.load()
    hist = list(hist)
    if any(h is None for h in hist):
        raise ValueError('solver_update: a history buffer is None (coefficients would shift onto the wrong buffers)')
    if len(hist) > 4 or len(coef) > 2 + 4 or (coef_dev is None and len(coef) > 2 and len(coef) - 2 != len(hist)):
        raise ValueError(f'solver_update: {len(coef)} coefficients for {len(hist)} history buffers')
    _chk_dev(out_x, xb, D, xs, out_m, thr, t_dev, coef_dev, *hist)
    cf = (C.c_float * 6)(*([float(c) for c in coef] + [0.0] * (6 - len(coef))))
    hp = (C.c_void_p * 4)(*([h.data_ptr() for h in hist] + [None] * (4 - len(hist))))
    B = xb.shape[0]
    n = xb[0].numel()
    if out_u8 is not None:
        if out_u8.dtype != torch.uint8 or out_u8.device != xb.device or not out_u8.is_contiguous() or xb.dim() != 4 or \
                tuple(out_u8.shape) != (B, xb.shape[2], xb.shape[3], xb.shape[1]):
            raise _lib.DsError('solver_update: out_u8 must be a contiguous uint8 [B, H, W, C] tensor on the same device')
        rc = lib.ds_solver_update_u8(_ptr(out_x), _ptr(out_m), out_u8.data_ptr(), xb.shape[1], xb.shape[2] * xb.shape[3], _ptr(xb), _ptr(xs),
                                     _ptr(D), hp, len(hist), _ptr(thr), int(mode), float(t), _ptr(t_dev), cf, _ptr(coef_dev), n, B, _stream(xb))
        _lib.check(rc, 'ds_solver_update_u8')
        LAUNCHES[0] += 1
        return out_x
    rc = lib.ds_solver_update(_ptr(out_x), _ptr(out_m), _ptr(xb), _ptr(xs), _ptr(D), hp, len(hist), _ptr(thr), int(mode), float(t),
                              _ptr(t_dev), cf, _ptr(coef_dev), n, B, _stream(xb))
    _lib.check(rc, 'ds_solver_update')
    LAUNCHES[0] += 1
    return out_x


def dyn_threshold(x0, q=0.995, floor=1.0, out=None):
    """Per-sample s = max(quantile(|x0|, q), floor) -> fp32 [B]."""
    lib = _lib.load()
    _chk_dev(x0)
    B = x0.shape[0]
    if out is None:
        out = torch.empty(B, device=x0.device, dtype=torch.float32)
    _lib.check(lib.ds_dyn_threshold(x0.data_ptr(), out.data_ptr(), B, x0[0].numel(), float(q), float(floor), _stream(x0)),
               'ds_dyn_threshold')
    LAUNCHES[0] += 1
    return out


def dynamic_thresholding_fn(x0):
    """The dynamic thresholding method (reference: solver_utils.py:77-86): clamp(x0, -s, s) / s with the per-sample
    0.995-quantile s (floored at 1)."""
    x0 = x0.contiguous().float()
    s = dyn_threshold(x0)
    out = torch.empty_like(x0)
    solver_update(None, x0, [0.0, 0.0], mode=S.DS_M_X0, D=x0, thr=s, out_m=out)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# DPM-Solver++ as linear-combination coefficients (reference: solver_utils.py:90-163, amed-solver-main/solver_utils.py:90-160)

def _lincomb(fn, nbasis):
    """Evaluate a formula that is linear in its tensor arguments on unit vectors to read off its coefficients."""
    eye = np.eye(nbasis)
    return fn(*[eye[k] for k in range(nbasis)])


def dpm_pp_coefs(t_prev, t, order, predict_x0=True, scale=1.0, xp=math):
    """Coefficients (cx, c_m0, c_m1, c_m2) such that the DPM-Solver++ multistep update of the given order is
    cx*x + c_m0*m0 + c_m1*m_{-1} + c_m2*m_{-2}.  t_prev = [.., t_{-2}, t_{-1}, t_0]; scalars (math) or torch tensors (xp=torch)."""
    lam = lambda s: -1 * xp.log(s)
    t0 = t_prev[-1]
    h = lam(t) - lam(t0)
    phi_1 = xp.expm1(-h) if predict_x0 else xp.expm1(h)
    zero = h * 0

    def formula(x, m0, m1, m2):
        if order == 1:
            return (t / t0) * x - scale * phi_1 * m0 if predict_x0 else x - scale * t * phi_1 * m0
        r0 = (lam(t0) - lam(t_prev[-2])) / h
        D1_0 = (1. / r0) * (m0 - m1)
        if order == 2:
            if predict_x0:
                return (t / t0) * x - scale * (phi_1 * m0 + 0.5 * phi_1 * D1_0)
            return x - scale * (t * phi_1 * m0 + 0.5 * t * phi_1 * D1_0)
        r1 = (lam(t_prev[-2]) - lam(t_prev[-3])) / h
        D1_1 = (1. / r1) * (m1 - m2)
        D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
        D2 = (1. / (r0 + r1)) * (D1_0 - D1_1)
        phi_2 = phi_1 / h + 1. if predict_x0 else phi_1 / h - 1.
        phi_3 = phi_2 / h - 0.5
        if predict_x0:
            return (t / t0) * x - scale * (phi_1 * m0 - phi_2 * D1 + phi_3 * D2)
        return x - scale * (t * phi_1 * m0 + t * phi_2 * D1 + t * phi_3 * D2)

    if order not in (1, 2, 3):
        raise ValueError("Solver order must be 1 or 2 or 3, got {}".format(order))
    # read the coefficients off by linearity: feed symbolic unit "tensors" (tuples of 4 coefficients)
    class V:
        def __init__(self, c):
            self.c = c

        def __add__(self, o):
            return V([a + b for a, b in zip(self.c, o.c)])

        def __sub__(self, o):
            return V([a - b for a, b in zip(self.c, o.c)])

        def __mul__(self, k):
            return V([a * k for a in self.c])

        __rmul__ = __mul__
    one = zero + 1
    basis = [V([one if j == k else zero for j in range(4)]) for k in range(4)]
    return formula(*basis).c


def dpm_pp_update(x, model_prev_list, t_prev_list, t, order, predict_x0=True, scale=1):
    """Tensor-in/tensor-out form kept for API compatibility with the reference (solver_utils.py:90-98); runs the fused kernel."""
    per_sample = any(torch.is_tensor(v) and v.numel() > 1 for v in list(t_prev_list) + [t, scale])
    if per_sample:
        B = x.shape[0]
        f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=x.device).reshape(-1).expand(B)
        c = dpm_pp_coefs([f(v) for v in t_prev_list], f(t), order, predict_x0, f(scale), xp=torch)
        cd = torch.zeros(6, B, device=x.device)
        cd[0], cd[2] = c[0], c[1]
        for k in range(1, order):
            cd[2 + k] = c[1 + k]
        hist = [model_prev_list[-1 - k] for k in range(order)]
        out = torch.empty_like(x)
        return solver_update(out, x.contiguous(), [0] * 6, mode=S.DS_M_NONE, hist=hist, coef_dev=cd.contiguous())
    fl = lambda v: float(v)
    c = dpm_pp_coefs([fl(v) for v in t_prev_list], fl(t), order, predict_x0, fl(scale))
    hist = [model_prev_list[-1 - k].contiguous() for k in range(order)]
    out = torch.empty_like(x)
    return solver_update(out, x.contiguous(), [c[0], 0.0] + list(c[1:1 + order]), mode=S.DS_M_NONE, hist=hist)


def dpm_solver_first_update(x, s, t, model_s=None, predict_x0=True, scale=1):
    return dpm_pp_update(x, [model_s], [s], t, 1, predict_x0=predict_x0, scale=scale)


def multistep_dpm_solver_second_update(x, model_prev_list, t_prev_list, t, predict_x0=True, scale=1):
    return dpm_pp_update(x, model_prev_list, t_prev_list, t, 2, predict_x0=predict_x0, scale=scale)


def multistep_dpm_solver_third_update(x, model_prev_list, t_prev_list, t, predict_x0=True, scale=1):
    return dpm_pp_update(x, model_prev_list, t_prev_list, t, 3, predict_x0=predict_x0, scale=scale)


# ---------------------------------------------------------------------------------------------------------------------
# UniPC (reference: solver_utils.py:174-287)

def unipc_coefs(t_prev, t, order, variant='bh1', predict_x0=True, use_corrector=True):
    """Host-side (float64) UniPC-p coefficients.  Returns (pred, corr) where
       pred = [cx, c_m0, c_m1, c_m2]               x_pred = cx*x + sum c_mk * m_{-k}
       corr = [cx, c_mt, c_m0, c_m1, c_m2] or None x_corr = cx*x + c_mt*model_t + sum c_mk * m_{-k}
    (t multiplies the history terms in eps-mode).  The <=3x3 systems are solved on the host: they depend only on t_steps."""
    assert order <= len(t_prev)
    lam = lambda s: -math.log(s)
    t0 = t_prev[-1]
    h = lam(t) - lam(t0)
    rks = [(lam(t_prev[-(i + 1)]) - lam(t0)) / h for i in range(1, order)] + [1.0]
    rks = np.array(rks, dtype=np.float64)
    hh = -h if predict_x0 else h
    h_phi_1 = math.expm1(hh)
    h_phi_k = h_phi_1 / hh - 1
    if variant == 'bh1':
        B_h = hh
    elif variant == 'bh2':
        B_h = math.expm1(hh)
    else:
        raise NotImplementedError()
    Rm, bv, fact = [], [], 1
    for i in range(1, order + 1):
        Rm.append(rks ** (i - 1))
        bv.append(h_phi_k * fact / B_h)
        fact *= (i + 1)
        h_phi_k = h_phi_k / hh - 1 / fact
    Rm, bv = np.stack(Rm), np.array(bv)
    nh = order - 1
    rhos_p = None
    if nh > 0:
        rhos_p = np.array([0.5]) if order == 2 else np.linalg.solve(Rm[:-1, :-1], bv[:-1])
    rhos_c = None
    if use_corrector:
        rhos_c = np.array([0.5]) if order == 1 else np.linalg.solve(Rm, bv)
    # base: x0-mode  t/t0*x - h_phi_1*m0 ; eps-mode  x - t*h_phi_1*m0 ; residual multiplier: B_h (x0) or t*B_h (eps)
    cx = t / t0 if predict_x0 else 1.0
    base_m0 = -h_phi_1 if predict_x0 else -t * h_phi_1
    mult = B_h if predict_x0 else t * B_h
    pred = [cx, base_m0, 0.0, 0.0]
    for k in range(nh):                 # D1s[k] = (m_{-(k+1)} - m0) / rk
        w = -mult * rhos_p[k] / rks[k]
        pred[2 + k] += w
        pred[1] -= w
    corr = None
    if use_corrector:
        corr = [cx, -mult * rhos_c[-1], base_m0 + mult * rhos_c[-1], 0.0, 0.0]
        for k in range(nh):
            w = -mult * rhos_c[k] / rks[k]
            corr[3 + k] += w
            corr[2] -= w
    return pred, corr


def unipc_update(x, model_prev_list, t_prev_list, t, order, x_t=None, variant='bh1', predict_x0=True, net=None, class_labels=None,
                 use_corrector=True):
    """Tensor-level UniPC step with the reference's signature (solver_utils.py:174-177): predictor kernel -> net -> corrector kernel."""
    assert order <= len(model_prev_list)
    tf = float(t)
    pred, corr = unipc_coefs([float(v) for v in t_prev_list], tf, order, variant, predict_x0, use_corrector)
    hist = [model_prev_list[-1 - k].contiguous() for k in range(order)]
    x = x.contiguous()
    if x_t is None:
        x_t = torch.empty_like(x)
        solver_update(x_t, x, [pred[0], 0.0] + pred[1:1 + order], mode=S.DS_M_NONE, hist=hist)
    model_t = None
    if use_corrector:
        t_dev = torch.as_tensor(t, dtype=torch.float32, device=x.device).reshape(1,)
        den = net(x_t, t_dev, class_labels)
        model_t = torch.empty_like(x)
        out = torch.empty_like(x)
        if predict_x0:
            s = dyn_threshold(den)
            solver_update(out, x, [corr[0], corr[1]] + corr[2:2 + order], mode=S.DS_M_X0, D=den, thr=s, hist=hist, out_m=model_t)
        else:
            solver_update(out, x, [corr[0], corr[1]] + corr[2:2 + order], mode=S.DS_M_EPS, D=den, xs=x_t, t=tf, hist=hist, out_m=model_t)
        x_t = out
    return x_t, model_t


# ---------------------------------------------------------------------------------------------------------------------
# DEIS coefficient tables (reference: solver_utils.py:297-400).  Host-side, once per schedule.

def edm2t(edm_steps, epsilon_s=1e-3, sigma_min=0.002, sigma_max=80):
    smin, smax = torch.tensor(sigma_min).cpu(), torch.tensor(sigma_max).cpu()
    beta_d = 2 * (np.log(smin ** 2 + 1) / epsilon_s - np.log(smax ** 2 + 1)) / (epsilon_s - 1)
    beta_min = np.log(smax ** 2 + 1) - 0.5 * beta_d
    sig = edm_steps.clone().detach().cpu()
    t_steps = ((beta_min ** 2 + 2 * beta_d * (sig ** 2 + 1).log()).sqrt() - beta_min) / beta_d
    return t_steps, beta_min, beta_d + beta_min


def cal_poly(prev_t, j, taus):
    poly = 1
    for k in range(prev_t.shape[0]):
        if k != j:
            poly = poly * (taus - prev_t[k]) / (prev_t[j] - prev_t[k])
    return poly


def t2alpha_fn(beta_0, beta_1, t):
    return torch.exp(-0.5 * t ** 2 * (beta_1 - beta_0) - t * beta_0)


def cal_intergrand(beta_0, beta_1, taus):
    with torch.enable_grad():
        taus.requires_grad_(True)
        alpha = t2alpha_fn(beta_0, beta_1, taus)
        alpha.log().sum().backward()
        d_log_alpha_dtau = taus.grad
    return -0.5 * d_log_alpha_dtau / torch.sqrt(alpha * (1 - alpha))


def get_deis_coeff_list(t_steps, max_order, N=10000, deis_mode='tab'):
    """Coefficient lists C[i] for deis_sampler ('tab': N-point quadrature of the Lagrange basis against the VP integrand;
    'rhoab': closed-form polynomial integrals in sigma)."""
    if deis_mode == 'tab':
        ts, beta_0, beta_1 = edm2t(t_steps)
        C_ = []
        for i, (t_cur, t_next) in enumerate(zip(ts[:-1], ts[1:])):
            order = min(i + 1, max_order)
            if order == 1:
                C_.append([])
                continue
            taus = torch.linspace(t_cur, t_next, N)
            dtau = (t_next - t_cur) / N
            prev_t = ts[[i - k for k in range(order)]]
            integrand = cal_intergrand(beta_0, beta_1, taus)
            C_.append([torch.sum(integrand * cal_poly(prev_t, j, taus)) * dtau for j in range(order)])
        return C_
    if deis_mode == 'rhoab':
        def quad2(a, b, s, e, c):
            return ((e ** 3 - s ** 3) / 3 - (e ** 2 - s ** 2) * (a + b) / 2 + (e - s) * a * b) / ((c - a) * (c - b))

        def quad3(a, b, c, s, e, d):
            num = (e ** 4 - s ** 4) / 4 - (e ** 3 - s ** 3) * (a + b + c) / 3 + (e ** 2 - s ** 2) * (a * b + a * c + b * c) / 2 \
                - (e - s) * a * b * c
            return num / ((d - a) * (d - b) * (d - c))
        C_, cur = [], None
        for i, (tc, tn) in enumerate(zip(t_steps[:-1], t_steps[1:])):
            order = min(i, max_order)
            if order == 0:
                C_.append([])
                continue
            p = t_steps[[i - k for k in range(order + 1)]]
            if order == 1:
                cur = [((tn - p[1]) ** 2 - (tc - p[1]) ** 2) / (2 * (tc - p[1])), (tn - tc) ** 2 / (2 * (p[1] - tc))]
            elif order == 2:
                cur = [quad2(p[1], p[2], tc, tn, tc), quad2(tc, p[2], tc, tn, p[1]), quad2(tc, p[1], tc, tn, p[2])]
            elif order == 3:
                cur = [quad3(p[1], p[2], p[3], tc, tn, tc), quad3(tc, p[2], p[3], tc, tn, p[1]),
                       quad3(tc, p[1], p[3], tc, tn, p[2]), quad3(tc, p[1], p[2], tc, tn, p[3])]
            # Reference quirk kept on purpose: rhoAB has no 4th-order formula (solver_utils.py:384-399), so from the fifth
            # step on (max_order == 4) the previous step's coefficients are reused.
            C_.append(cur)
        return C_
    return None

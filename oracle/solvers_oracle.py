"""ORACLE — test infrastructure, not product code (see oracle/edm_oracle.py header).

CPU restatement of the reference's solver loops and solver math, written as one scaffold
(`sample`) plus per-solver step rules.  Citations are into /root/reference/diff-solvers-main/
(solvers.py, solver_utils.py), /root/reference/amed-solver-main/ and /root/reference/gits-main/.
Pinned by tests/golden/solvers_*.npz (generated from the real reference by oracle/gen_golden.py).
"""
import numpy as np
import torch

# ------------------------------------------------------------------------------------------------
# schedules


def get_schedule(num_steps, sigma_min, sigma_max, device=None, schedule_type='polynomial', schedule_rho=7, net=None, dp_list=None):
    """solver_utils.py:6-52; the dp_list gather is gits-main/solver_utils.py:52-53."""
    if schedule_type == 'polynomial':                                     # :25-27
        i = torch.arange(num_steps, device=device)
        a, b = sigma_max ** (1 / schedule_rho), sigma_min ** (1 / schedule_rho)
        t = (a + i / (num_steps - 1) * (b - a)) ** schedule_rho
    elif schedule_type == 'logsnr':                                       # :28-32
        lo = (-1 * torch.log(torch.tensor(sigma_max))).item()
        hi = (-1 * torch.log(torch.tensor(sigma_min))).item()
        t = (-torch.linspace(lo, hi, steps=num_steps, device=device)).exp()
    elif schedule_type == 'time_uniform':                                 # :33-41 (VP time-uniform, numpy scalars on CPU)
        eps_s = 1e-3
        smin, smax = torch.tensor(sigma_min), torch.tensor(sigma_max)
        beta_d = 2 * (np.log(smin ** 2 + 1) / eps_s - np.log(smax ** 2 + 1)) / (eps_s - 1)
        beta_min = np.log(smax ** 2 + 1) - 0.5 * beta_d
        i = torch.arange(num_steps, device=device)
        tt = (1 + i / (num_steps - 1) * (eps_s ** (1 / schedule_rho) - 1)) ** schedule_rho
        tt = tt.clone().detach().cpu()
        t = (np.e ** (0.5 * beta_d.clone() * (tt ** 2) + beta_min.clone() * tt) - 1) ** 0.5
    elif schedule_type == 'discrete':                                     # :42-48
        assert net is not None
        tmin = net.sigma_inv(torch.tensor(sigma_min, device=device))
        tmax = net.sigma_inv(torch.tensor(sigma_max, device=device))
        i = torch.arange(num_steps, device=device)
        t = net.sigma((tmax + i / (num_steps - 1) * (tmin ** (1 / schedule_rho) - tmax)) ** schedule_rho)
    else:
        raise ValueError('Got wrong schedule type {}'.format(schedule_type))
    if dp_list is not None:
        return t[dp_list].to(device)
    return t.to(device)


# ------------------------------------------------------------------------------------------------
# solver math


def dynamic_thresholding(x0, p=0.995):
    """solver_utils.py:77-86."""
    s = torch.quantile(x0.abs().reshape(x0.shape[0], -1), p, dim=1)
    s = torch.maximum(s, torch.ones_like(s)).reshape(-1, *([1] * (x0.dim() - 1)))
    return torch.clamp(x0, -s, s) / s


def quantile_by_sort(rows, q):
    """What torch.quantile(rows, q, dim=1) computes for fp32 input (ATen quantile_compute, 'linear'):
    rank = q*(n-1) in fp32, lerp between the two neighbouring order statistics."""
    n = rows.shape[1]
    srt = rows.sort(dim=1).values
    rank = torch.tensor(q, dtype=rows.dtype) * (n - 1)
    lo = rank.floor()
    w = rank - lo
    lo_i = int(lo.item())
    hi_i = min(int(rank.ceil().item()), n - 1)
    return torch.lerp(srt[:, lo_i], srt[:, hi_i], w)


def _r4(t):
    return t.reshape(-1, 1, 1, 1)


def dpm_pp_update(x, models, ts, t, order, predict_x0=True, scale=1):
    """solver_utils.py:90-163 (scale= argument: amed-solver-main/solver_utils.py:90-160)."""
    t = _r4(t)
    t0 = _r4(ts[-1])
    lam = lambda s: -1 * s.log()
    h = lam(t) - lam(t0)
    phi1 = torch.expm1(-h) if predict_x0 else torch.expm1(h)
    m0 = models[-1]
    if order == 1:                                                        # :102-113
        return (t / t0) * x - scale * phi1 * m0 if predict_x0 else x - scale * t * phi1 * m0
    t1 = _r4(ts[-2])
    r0 = (lam(t0) - lam(t1)) / h
    D1_0 = (1. / r0) * (m0 - models[-2])
    if order == 2:                                                        # :117-133
        if predict_x0:
            return (t / t0) * x - scale * (phi1 * m0 + 0.5 * phi1 * D1_0)
        return x - scale * (t * phi1 * m0 + 0.5 * t * phi1 * D1_0)
    if order == 3:                                                        # :137-163
        t2 = _r4(ts[-3])
        r1 = (lam(t1) - lam(t2)) / h
        D1_1 = (1. / r1) * (models[-2] - models[-3])
        D1 = D1_0 + (r0 / (r0 + r1)) * (D1_0 - D1_1)
        D2 = (1. / (r0 + r1)) * (D1_0 - D1_1)
        phi2 = phi1 / h + 1. if predict_x0 else phi1 / h - 1.
        phi3 = phi2 / h - 0.5
        if predict_x0:
            return (t / t0) * x - scale * (phi1 * m0 - phi2 * D1 + phi3 * D2)
        return x - scale * (t * phi1 * m0 + t * phi2 * D1 + t * phi3 * D2)
    raise ValueError('Solver order must be 1 or 2 or 3, got {}'.format(order))


def unipc_update(x, models, ts, t, order, variant='bh1', predict_x0=True, net=None, class_labels=None, use_corrector=True):
    """solver_utils.py:174-287."""
    assert order <= len(models)
    t0 = ts[-1].reshape(1,)
    t = t.reshape(1,)
    lam0, lamt = -1 * t0.log(), -1 * t.log()
    m0 = models[-1]
    h = lamt - lam0
    rks, D1s = [], []
    for i in range(1, order):                                             # :191-197
        rk = ((-1 * ts[-(i + 1)].reshape(1,).log()) - lam0) / h
        rks.append(rk)
        D1s.append((models[-(i + 1)] - m0) / rk)
    rks.append(1.)
    rks = torch.tensor(rks, device=x.device)
    hh = -h if predict_x0 else h
    h_phi_1 = torch.expm1(hh)
    h_phi_k = h_phi_1 / hh - 1
    if variant == 'bh1':
        B_h = hh
    elif variant == 'bh2':
        B_h = torch.expm1(hh)
    else:
        raise NotImplementedError()
    R, b, fact = [], [], 1
    for i in range(1, order + 1):                                         # :218-222
        R.append(torch.pow(rks, i - 1))
        b.append(h_phi_k * fact / B_h)
        fact *= (i + 1)
        h_phi_k = h_phi_k / hh - 1 / fact
    R, b = torch.stack(R), torch.cat(b)
    have_hist = len(D1s) > 0
    if have_hist:
        D1s = torch.stack(D1s, dim=1)
        rhos_p = torch.tensor([0.5], device=b.device) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
    if use_corrector:
        rhos_c = torch.tensor([0.5], device=b.device) if order == 1 else torch.linalg.solve(R, b)
    mix = lambda rho: torch.einsum('k,bkchw->bchw', rho, D1s)
    model_t = None
    if predict_x0:                                                        # :250-267
        base = t / t0 * x - h_phi_1 * m0
        x_t = base - B_h * (mix(rhos_p) if have_hist else 0)
        if use_corrector:
            model_t = dynamic_thresholding(net(x_t, t, class_labels))
            corr = mix(rhos_c[:-1]) if have_hist else 0
            x_t = base - B_h * (corr + rhos_c[-1] * (model_t - m0))
    else:                                                                 # :268-285
        base = x - t * h_phi_1 * m0
        x_t = base - t * B_h * (mix(rhos_p) if have_hist else 0)
        if use_corrector:
            model_t = (x_t - net(x_t, t, class_labels)) / t
            corr = mix(rhos_c[:-1]) if have_hist else 0
            x_t = base - t * B_h * (corr + rhos_c[-1] * (model_t - m0))
    return x_t, model_t


def get_deis_coeff_list(t_steps, max_order, N=10000, deis_mode='tab'):
    """solver_utils.py:297-400 (edm2t :297-303, cal_poly :307-313, cal_intergrand :323-331)."""
    if deis_mode == 'tab':
        eps_s, smin, smax = 1e-3, torch.tensor(0.002), torch.tensor(80)
        beta_d = 2 * (np.log(smin ** 2 + 1) / eps_s - np.log(smax ** 2 + 1)) / (eps_s - 1)
        beta_0 = np.log(smax ** 2 + 1) - 0.5 * beta_d
        beta_1 = beta_d + beta_0
        sig = t_steps.clone().detach().cpu()
        ts = ((beta_0 ** 2 + 2 * beta_d * (sig ** 2 + 1).log()).sqrt() - beta_0) / beta_d
        C = []
        for i, (t_cur, t_next) in enumerate(zip(ts[:-1], ts[1:])):
            order = min(i + 1, max_order)
            if order == 1:
                C.append([])
                continue
            taus = torch.linspace(t_cur, t_next, N)
            dtau = (t_next - t_cur) / N
            prev_t = ts[[i - k for k in range(order)]]
            with torch.enable_grad():
                taus.requires_grad_(True)
                alpha = torch.exp(-0.5 * taus ** 2 * (beta_1 - beta_0) - taus * beta_0)
                alpha.log().sum().backward()
                dlog = taus.grad
            integrand = -0.5 * dlog / torch.sqrt(alpha * (1 - alpha))
            row = []
            for j in range(order):
                poly = 1
                for k in range(order):
                    if k != j:
                        poly = poly * (taus - prev_t[k]) / (prev_t[j] - prev_t[k])
                row.append(torch.sum(integrand * poly) * dtau)
            C.append(row)
        return C
    if deis_mode == 'rhoab':
        def I2(a, b, s, e, c):                                            # :367-369
            return ((e ** 3 - s ** 3) / 3 - (e ** 2 - s ** 2) * (a + b) / 2 + (e - s) * a * b) / ((c - a) * (c - b))

        def I3(a, b, c, s, e, d):                                         # :372-375
            co = (e ** 4 - s ** 4) / 4 - (e ** 3 - s ** 3) * (a + b + c) / 3 + (e ** 2 - s ** 2) * (a * b + a * c + b * c) / 2 - (e - s) * a * b * c
            return co / ((d - a) * (d - b) * (d - c))
        C, row = [], None
        for i, (tc, tn) in enumerate(zip(t_steps[:-1], t_steps[1:])):
            order = min(i, max_order)
            if order == 0:
                C.append([])
                continue
            p = t_steps[[i - k for k in range(order + 1)]]
            if order == 1:
                row = [((tn - p[1]) ** 2 - (tc - p[1]) ** 2) / (2 * (tc - p[1])), (tn - tc) ** 2 / (2 * (p[1] - tc))]
            elif order == 2:
                row = [I2(p[1], p[2], tc, tn, tc), I2(tc, p[2], tc, tn, p[1]), I2(tc, p[1], tc, tn, p[2])]
            elif order == 3:
                row = [I3(p[1], p[2], p[3], tc, tn, tc), I3(tc, p[2], p[3], tc, tn, p[1]), I3(tc, p[1], p[3], tc, tn, p[2]),
                       I3(tc, p[1], p[2], tc, tn, p[3])]
            # order >= 4 has no branch in the reference (:384-399): the previous step's row is appended again
            C.append(row)
        return C
    return None


def _abv_coeffs(t_steps, i, order):
    """Variable-step Adams-Bashforth coefficients, solvers.py:451-477."""
    h = lambda a: t_steps[a + 1] - t_steps[a]
    hn, h1 = h(i), h(i - 1)
    if order == 2:
        return [(2 + hn / h1) / 2, -(hn / h1) / 2]
    h2 = h(i - 2)
    tA = (1 - hn / (3 * (hn + h1)) * (hn * (hn + h1)) / (h1 * (h1 + h2))) / 2
    if order == 3:
        return [(2 + hn / h1) / 2 + tA, -(hn / h1) / 2 - (1 + h1 / h2) * tA, tA * h1 / h2]
    h3 = h(i - 3)
    tB = ((1 - hn / (3 * (hn + h1))) / 2 + (1 - hn / (2 * (hn + h1))) * hn / (6 * (hn + h1 + h2))) \
        * (hn * (hn + h1) * (hn + h1 + h2)) / (h1 * (h1 + h2) * (h1 + h2 + h3))
    g = h1 * (h1 + h2) / (h2 * (h2 + h3))
    return [(2 + hn / h1) / 2 + tA + tB,
            -(hn / h1) / 2 - (1 + h1 / h2) * tA - (1 + h1 / h2 + g) * tB,
            tA * h1 / h2 + (h1 / h2 + g * (1 + h2 / h3)) * tB,
            -tB * g * h1 / h2]


# ------------------------------------------------------------------------------------------------
# the sampling scaffold (solvers.py:63-96 and the identical prologue/epilogue of every sampler)


def _denoise(net, x, t, class_labels, condition, unconditional_condition):
    """solvers.py:9-14 get_denoised."""
    if hasattr(net, 'guidance_type'):
        return net(x, t, condition=condition, unconditional_condition=unconditional_condition)
    return net(x, t, class_labels=class_labels)


@torch.no_grad()
def sample(net, latents, solver, class_labels=None, condition=None, unconditional_condition=None, num_steps=None, sigma_min=0.002,
           sigma_max=80, schedule_type='polynomial', schedule_rho=7, afs=False, denoise_to_zero=False, return_inters=False,
           return_eps=False, t_steps=None, max_order=None, r=0.5, coeff_list=None, predict_x0=True, lower_order_final=True,
           variant='bh2', **_):
    """One scaffold for euler (solvers.py:18-96), heun (:100-183), dpm_2 (:187-273), ipndm (:277-374), ipndm_v (:378-499),
    deis (:503-607), dpm_pp (:612-713) and unipc (:717-821)."""
    if t_steps is None:
        t_steps = get_schedule(num_steps, sigma_min, sigma_max, device=latents.device, schedule_type=schedule_type,
                               schedule_rho=schedule_rho, net=net)
    D = lambda x, t: _denoise(net, x, t, class_labels, condition, unconditional_condition)
    if max_order is None:
        max_order = 3 if solver in ('dpm_pp', 'unipc') else 4
    x_next = latents * t_steps[0]
    inters, eps_hist = [x_next.unsqueeze(0)], []
    hist, hist_t = [], []

    if solver == 'unipc':                                                 # :780-811
        assert 0 < max_order < 4
        if afs:
            d0 = x_next / ((1 + t_steps[0] ** 2).sqrt())
            den = x_next - t_steps[0] * d0
        else:
            den = D(x_next, t_steps[0])
            d0 = (x_next - den) / t_steps[0]
        hist, hist_t = [dynamic_thresholding(den) if predict_x0 else d0], [t_steps[0]]
        for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
            x_cur = x_next
            if i + 1 < max_order:
                x_next, m = unipc_update(x_cur, hist, hist_t, t_next, i + 1, net=net, class_labels=class_labels, use_corrector=True,
                                         predict_x0=predict_x0, variant=variant)
                hist.append(m)
                hist_t.append(t_next)
            else:
                order = min(max_order, num_steps - i - 1) if lower_order_final else max_order
                x_next, m = unipc_update(x_cur, hist, hist_t, t_next, order, net=net, class_labels=class_labels,
                                         use_corrector=(i != num_steps - 2), predict_x0=predict_x0, variant=variant)
                for k in range(max_order - 1):
                    hist[k], hist_t[k] = hist[k + 1], hist_t[k + 1]
                hist_t[-1] = t_next
                if i < num_steps - 2:
                    hist[-1] = m
            if return_inters:
                inters.append(x_next.unsqueeze(0))
    else:
        if solver in ('ipndm', 'ipndm_v', 'deis'):
            assert 1 <= max_order <= 4
        if solver == 'deis':
            assert coeff_list is not None
        if solver == 'dpm_pp':
            assert 1 <= max_order <= 3
        for i, (t_cur, t_next) in enumerate(zip(t_steps[:-1], t_steps[1:])):
            x_cur = x_next
            first = (i == 0) if solver in ('euler', 'heun', 'dpm_2', 'ipndm', 'dpm_pp') else (len(hist) == 0)
            if afs and first:                                             # :76-77 analytical first step
                d_cur = x_cur / ((1 + t_cur ** 2).sqrt())
                den = x_cur - t_cur * d_cur
            else:
                den = D(x_cur, t_cur)
                d_cur = (x_cur - den) / t_cur
            hstep = t_next - t_cur
            if solver == 'euler':                                         # :81
                x_next = x_cur + hstep * d_cur
            elif solver == 'heun':                                        # :163-168
                x_next = x_cur + hstep * d_cur
                d_prime = (x_next - D(x_next, t_next)) / t_next
                x_next = x_cur + hstep * (0.5 * d_cur + 0.5 * d_prime)
            elif solver == 'dpm_2':                                       # :252-258
                t_mid = (t_next ** r) * (t_cur ** (1 - r))
                x_next = x_cur + (t_mid - t_cur) * d_cur
                d_prime = (x_next - D(x_next, t_mid)) / t_mid
                x_next = x_cur + hstep * ((1 / (2 * r)) * d_prime + (1 - 1 / (2 * r)) * d_cur)
            elif solver in ('ipndm', 'ipndm_v', 'deis'):
                order = min(max_order, i + 1)
                prev = hist[::-1]                                         # prev[0] is the most recent stored d
                if order == 1:
                    x_next = x_cur + hstep * d_cur
                elif solver == 'ipndm':                                   # :346-352
                    if order == 2:
                        x_next = x_cur + hstep * (3 * d_cur - prev[0]) / 2
                    elif order == 3:
                        x_next = x_cur + hstep * (23 * d_cur - 16 * prev[0] + 5 * prev[1]) / 12
                    else:
                        x_next = x_cur + hstep * (55 * d_cur - 59 * prev[0] + 37 * prev[1] - 9 * prev[2]) / 24
                elif solver == 'ipndm_v':                                 # :451-477
                    c = _abv_coeffs(t_steps, i, order)
                    acc = c[0] * d_cur
                    for k in range(1, order):
                        acc = acc + c[k] * prev[k - 1]
                    x_next = x_cur + hstep * acc
                else:                                                     # deis :576-585
                    c = coeff_list[i]
                    x_next = x_cur + c[0] * d_cur
                    for k in range(1, order):
                        x_next = x_next + c[k] * prev[k - 1]
                if len(hist) == max_order - 1:                            # :358-363 (indexes an empty list when max_order == 1, as the reference does)
                    for k in range(max_order - 2):
                        hist[k] = hist[k + 1]
                    hist[-1] = d_cur
                else:
                    hist.append(d_cur)
            elif solver == 'dpm_pp':                                      # :685-702
                hist.append(dynamic_thresholding(den) if predict_x0 else d_cur)
                hist_t.append(t_cur)
                if lower_order_final:
                    order = i + 1 if i + 1 < max_order else min(max_order, num_steps - (i + 1))
                else:
                    order = min(max_order, i + 1)
                x_next = dpm_pp_update(x_cur, hist, hist_t, t_next, order, predict_x0=predict_x0)
                hist, hist_t = hist[-3:], hist_t[-3:]
            else:
                raise NotImplementedError(solver)
            if return_inters:
                inters.append(x_next.unsqueeze(0))
            if return_eps:
                eps_hist.append(d_cur.unsqueeze(0))
    if denoise_to_zero:                                                   # :87-90
        x_next = D(x_next, t_next)
        if return_inters:
            inters.append(x_next.unsqueeze(0))
    if return_inters:
        if return_eps and solver != 'unipc':
            return torch.cat(inters, dim=0), torch.cat(eps_hist, dim=0)
        return torch.cat(inters, dim=0)
    return x_next


# ------------------------------------------------------------------------------------------------
# GITS (gits-main/gits_utils.py)


def dp(cost_mat, num_steps, num_steps_tea, coeff):
    """gits_utils.py:185-203: float64 dynamic programme and exact-equality back-trace."""
    K = num_steps - 1
    V = np.full((num_steps_tea, K + 1), np.inf)
    for i in range(num_steps_tea):
        V[i][1] = cost_mat[i][-1]
    for k in range(2, K + 1):
        for j in range(num_steps_tea - 1):
            for i in range(j + 1, num_steps_tea - 1):
                V[j][k] = min(V[j][k], cost_mat[j][i] + coeff * V[i][k - 1])
    phi, w = [0], 0
    for temp in range(K):
        k = K - temp
        for j in range(w + 1, num_steps_tea):
            if V[w][k] == cost_mat[w][j] + coeff * V[j][k - 1]:
                phi.append(j)
                w = j
                break
    phi.append(num_steps_tea - 1)
    return phi


def cal_deviation(traj, ch, r, bs=1):
    """gits_utils.py:237-255: distance of each intermediate point from the chord start->end."""
    traj = traj.transpose(0, 1)
    a, b, c = traj[:, 1:-1], traj[:, 0].unsqueeze(1), traj[:, -1].unsqueeze(1)
    ac, bc = c - a, c - b
    unit = bc / torch.norm(bc, p=2, dim=(1, 2, 3, 4)).reshape(bs, 1, 1, 1, 1)
    proj = torch.sum(ac * unit.expand_as(ac), dim=(2, 3, 4))[:, :, None, None, None] * unit
    return torch.norm(ac - proj, p=2, dim=(2, 3, 4))


def gits_cost_matrix(teacher_traj, eps_traj, t_steps, metric, ch, res):
    """gits_utils.py:110-132: cost of a single Euler jump i -> j measured against the teacher trajectory."""
    n = t_steps.shape[0]
    bs = teacher_traj.shape[1]
    cost = torch.zeros((n, n))
    dev_tea = cal_deviation(teacher_traj, ch, res, bs=bs).mean(dim=0)
    dev_tea = torch.cat([dev_tea, torch.zeros_like(dev_tea[:1])])
    for i in range(n - 1):
        for j in range(i + 1, n):
            x_next = teacher_traj[i] + (t_steps[j] - t_steps[i]) * eps_traj[i]
            if metric == 'l1':
                cost[i][j] += torch.norm(x_next - teacher_traj[j], p=1, dim=(1, 2, 3)).mean()
            elif metric == 'l2':
                cost[i][j] += torch.norm(x_next - teacher_traj[j], p=2, dim=(1, 2, 3)).mean()
            elif metric == 'dev':
                temp = torch.cat((teacher_traj[0].unsqueeze(0), x_next.unsqueeze(0), teacher_traj[-1].unsqueeze(0)), dim=0)
                cost[i][j] += (cal_deviation(temp, ch, res, bs=bs).mean(dim=0) - dev_tea[j - 1]).mean()
            else:
                raise NotImplementedError(metric)
    return cost

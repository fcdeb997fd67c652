"""gits_utils — drop-in for gits-main/gits_utils.py: search a coarse time schedule (an integer index list into the fine
teacher grid) by dynamic programming over the cost of single Euler jumps along teacher trajectories.

Reference: gits_utils.py:15-37 (get_sampler_fn), :42-180 (get_dp_list), :185-232 (dp), :237-255 (cal_deviation).
The teacher trajectory comes from the native samplers (`return_inters=True, return_eps=True`) and stays on the device; the
O(N_tea^2) loop of tiny reductions (:115-132) is ONE kernel (`ds_gits_cost`).  The DP itself is the reference's float64 numpy
recurrence with its exact-equality back-trace (:190-203) — the index list must be bit-identical for a given cost matrix.
"""
import copy
import ctypes as C

import numpy as np
import torch

from . import _lib
from . import solver_utils, solvers


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank(), dist
    return 1, 0, None


def get_sampler_fn(solver, device, dp_list=None, net=None, **kwargs):
    """Reference: gits_utils.py:15-37."""
    table = {'euler': solvers.euler_sampler, 'heun': solvers.heun_sampler, 'dpm': solvers.dpm_2_sampler, 'ipndm': solvers.ipndm_sampler,
             'ipndm_v': solvers.ipndm_v_sampler, 'dpmpp': solvers.dpm_pp_sampler, 'deis': solvers.deis_sampler}
    if solver not in table:
        raise NotImplementedError(f"Unknown solver: {solver}")
    if solver == 'deis':
        t_steps = solver_utils.get_schedule(kwargs['num_steps_tea'], kwargs['sigma_min'], kwargs['sigma_max'], device=device,
                                            schedule_type=kwargs["schedule_type"], schedule_rho=kwargs["schedule_rho"], net=net, dp_list=dp_list)
        return table[solver], solver_utils.get_deis_coeff_list(t_steps, kwargs['max_order'], deis_mode=kwargs["deis_mode"])
    return table[solver], None


def cal_deviation(traj, ch, r, bs=1):
    """Deviation of every intermediate trajectory point from the chord start -> end.  Reference: gits_utils.py:237-255."""
    traj = traj.transpose(0, 1)
    a, b, c = traj[:, 1:-1], traj[:, 0].unsqueeze(1), traj[:, -1].unsqueeze(1)
    ac, bc = c - a, c - b
    unit = bc / torch.norm(bc, p=2, dim=(1, 2, 3, 4)).reshape(bs, 1, 1, 1, 1)
    coef = torch.sum(ac * unit.expand_as(ac), dim=(2, 3, 4))
    perp = ac - coef[:, :, None, None, None] * unit
    return torch.norm(perp, p=2, dim=(2, 3, 4))


def cost_matrix(teacher_traj, eps_traj, t_steps, metric, ch, res):
    """cost[i][j] of replacing the teacher's path i -> j by one Euler jump (gits_utils.py:110-132), batch-averaged.
    One kernel launch for all pairs; the few remaining scalar steps are tiny torch ops on the device."""
    lib = _lib.load()
    N, B = teacher_traj.shape[0], teacher_traj.shape[1]
    n = teacher_traj[0, 0].numel()
    traj = teacher_traj.to(torch.float32).contiguous()
    eps = eps_traj.to(torch.float32).contiguous()
    t = t_steps.to(device=traj.device, dtype=torch.float32).contiguous()
    out = torch.zeros(N, N, B, 4, dtype=torch.float64, device=traj.device)
    stream = C.c_void_p(torch.cuda.current_stream(traj.device).cuda_stream)
    _lib.check(lib.ds_gits_cost(traj.data_ptr(), eps.data_ptr(), t.data_ptr(), out.data_ptr(), N, B, n, stream), 'ds_gits_cost')
    solver_utils.LAUNCHES[0] += 1
    upper = torch.triu(torch.ones(N, N, device=traj.device), diagonal=1)
    if metric == 'l1':
        cost = out[..., 0].mean(dim=2)
    elif metric == 'l2':
        cost = out[..., 1].sqrt().mean(dim=2)
    elif metric == 'dev':
        bc2 = (traj[-1] - traj[0]).double().pow(2).sum(dim=(1, 2, 3))                      # |c - b|^2 per sample
        dev_stu = (out[..., 2] - out[..., 3] ** 2 / bc2).clamp_min(0).sqrt()              # |perp|: Pythagoras on the chord
        dev_tea = cal_deviation(traj, ch, res, bs=B).mean(dim=0)
        dev_tea = torch.cat([dev_tea, torch.zeros_like(dev_tea[:1])]).double()            # index j-1
        shift = torch.zeros(N, dtype=torch.float64, device=traj.device)
        shift[1:] = dev_tea[:N - 1]
        cost = (dev_stu - shift[None, :, None]).mean(dim=2)
    else:
        raise NotImplementedError(f"Unknown metric: {metric}")
    return (cost * upper).to(torch.float32)


def get_dp_list(net, device, **solver_kwargs):
    """Reference: gits_utils.py:42-180 (EDM-style nets; the ms_coco / LDM prompt branches are caller-side data plumbing and
    out of scope).  Returns the python list of teacher-grid indices."""
    kwargs = copy.deepcopy(solver_kwargs)
    num_warmup, max_batch_size = kwargs['num_warmup'], kwargs['max_batch_size']
    sigma_min, sigma_max = kwargs['sigma_min'], kwargs['sigma_max']
    num_steps, num_steps_tea = kwargs['num_steps'], kwargs['num_steps_tea']
    schedule_type, schedule_rho = kwargs['schedule_type'], kwargs['schedule_rho']
    afs, metric, coeff = kwargs['afs'], kwargs['metric'], kwargs['coeff']
    model_source = kwargs.get('model_source', 'edm')
    world, rank, dist = _dist()
    kwargs['solver'] = solver_kwargs['solver_tea']
    sampler_fn_tea, coeff_list = get_sampler_fn(device=device, net=net, dp_list=[i for i in range(num_steps_tea)], **kwargs)
    kwargs['t_steps'] = t_steps = solvers.get_schedule(num_steps_tea, sigma_min, sigma_max, device=device, schedule_type=schedule_type,
                                                       schedule_rho=schedule_rho, net=net)
    kwargs['coeff_list'] = coeff_list
    kwargs['return_inters'] = True
    kwargs['return_eps'] = True
    kwargs['num_steps'] = num_steps_tea
    rounds = num_warmup // (max_batch_size + 1) + 1
    batch_gpu = max_batch_size // world
    cost_mat = torch.zeros((num_steps_tea, num_steps_tea), device=device)
    for _ in range(rounds):
        latents = torch.randn([batch_gpu, net.img_channels, net.img_resolution, net.img_resolution], device=device)
        class_labels = None
        if net.label_dim:
            if model_source == 'adm':
                class_labels = torch.randint(net.label_dim, size=(batch_gpu,), device=device)
            else:
                class_labels = torch.eye(net.label_dim, device=device)[torch.randint(net.label_dim, size=[batch_gpu], device=device)]
        teacher_traj, eps_traj = sampler_fn_tea(net, latents, class_labels=class_labels, **kwargs)
        cost_mat += cost_matrix(teacher_traj, eps_traj, t_steps, metric, net.img_channels, net.img_resolution)
    if dist is not None:
        dist.all_reduce(cost_mat)
    cost_mat /= world * rounds
    cost_np = cost_mat.detach().cpu().numpy()
    dp_list = phi = dp(cost_np, num_steps, num_steps_tea, coeff, False, None, t_steps)
    kwargs['return_inters'] = False
    kwargs['return_eps'] = False
    kwargs['solver'] = solver_kwargs['solver']
    kwargs['num_steps'] = solver_kwargs['num_steps']
    if afs:
        dist_min = 999999
        for k in range(1, phi[1]):
            cand = copy.deepcopy(phi)
            cand.insert(1, k)
            sampler_fn, solver_kwargs['coeff_list'] = get_sampler_fn(device=device, dp_list=cand, **kwargs)
            kwargs['t_steps'] = solvers.get_schedule(num_steps_tea, sigma_min, sigma_max, device=device, schedule_type=schedule_type,
                                                     schedule_rho=schedule_rho, net=net, dp_list=cand)
            images_afs = sampler_fn(net, latents, class_labels=class_labels, **kwargs)
            dist_temp = torch.norm(images_afs - teacher_traj[-1], p=2, dim=(1, 2, 3)).mean()
            if dist is not None:
                dist.all_reduce(dist_temp)
            dist_temp /= world
            if dist_temp < dist_min:
                dist_min = dist_temp
                dp_list = cand
    return dp_list


def dp(cost_mat, num_steps, num_steps_tea, coeff, multiple_coeff=False, desc=None, t_steps=None):
    """Dynamic programme over the teacher grid.  Reference: gits_utils.py:185-232 (the multiple_coeff branch only writes a text
    report for MS-COCO sweeps and is not reproduced).  V[j][k] = min_i cost[j][i] + coeff*V[i][k-1] in float64; the back-trace
    selects the first j whose cost reproduces V exactly, as the reference does."""
    K = num_steps - 1
    V = np.full((num_steps_tea, K + 1), np.inf)
    V[:, 1] = cost_mat[:, -1]
    for k in range(2, K + 1):
        for j in range(num_steps_tea - 1):
            for i in range(j + 1, num_steps_tea - 1):
                V[j][k] = min(V[j][k], cost_mat[j][i] + coeff * V[i][k - 1])
    phi, w = [0], 0
    for step in range(K):
        k = K - step
        for j in range(w + 1, num_steps_tea):
            if V[w][k] == cost_mat[w][j] + coeff * V[j][k - 1]:
                phi.append(j)
                w = j
                break
    phi.append(num_steps_tea - 1)
    return phi

"""solvers_amed — drop-in for amed-solver-main/solvers_amed.py (AMED-Solver and the AMED plug-in variants of
Euler / iPNDM / DPM-Solver-2 / DPM-Solver++; plain Heun for teachers).

Same function names and argument lists as the reference (solvers_amed.py:69-708).  Per step the predictor supplies per-sample
(r, scale_dir, scale_time) of shape [B,1,1,1] (solvers_amed.py:22-55); here they stay on the device and enter the fused update
kernel as per-sample coefficient vectors, and the second network evaluation runs at per-sample sigma = scale_time * t_mid.
The U-Net bottleneck read-out (forward hook + channel mean, solvers_amed.py:7-27) is an output of the native denoiser
(`bottleneck=`), so no module hooks are needed.  Training (`train=True`, autograd through the predictor) is out of scope
(SURVEY.md section 8): the keyword arguments are accepted for signature compatibility and rejected at run time.
"""
import torch

from . import _cstructs as S
from . import solvers as base
from .ldm_net import B200LDMNet
from .net import B200Net
from .solver_utils import *                        # noqa: F401,F403
from .solver_utils import dpm_pp_coefs, dyn_threshold, get_schedule, solver_update
from .solvers import _Loop, _afs_div, as_native

AB = {1: [1.0], 2: [3 / 2, -1 / 2], 3: [23 / 12, -16 / 12, 5 / 12], 4: [55 / 24, -59 / 24, 37 / 24, -9 / 24]}


def get_denoised(net, x, t, class_labels=None, condition=None, unconditional_condition=None):
    return base.get_denoised(net, x, t, class_labels=class_labels, condition=condition, unconditional_condition=unconditional_condition)


def _no_train(train):
    if train:
        raise NotImplementedError('diff_sampler_b200.solvers_amed: the AMED training path (train=True) is out of scope; '
                                  'use the reference implementation for training the predictor')


class _Amed(_Loop):
    """Sampling loop state plus the predictor plumbing."""

    def __init__(self, predictor, net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
                 schedule_type, schedule_rho, return_inters, denoise_to_zero):
        super().__init__(net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max, schedule_type,
                         schedule_rho, None, return_inters, False, denoise_to_zero)
        self.pred = self._native_predictor(predictor)
        self.B = self.latents.shape[0]
        self.bott = torch.zeros(self.B, 64, device=self.latents.device)
        self.native = isinstance(self.net, B200Net)
        self.native_ldm = isinstance(self.net, B200LDMNet)

    @staticmethod
    def _native_predictor(predictor):
        """The predictor as an AMEDPredictor (one kernel per step) when it is one, or the reference's AMED_predictor module (converted once
        and cached on it); any other callable is used as it is through its torch forward."""
        from .amed_predictor import AMEDPredictor
        if isinstance(predictor, AMEDPredictor):
            return predictor
        inner = getattr(predictor, 'module', predictor)
        if type(inner).__name__ == 'AMED_predictor' and hasattr(inner, 'state_dict'):
            fp = base._weights_fingerprint(inner)
            hit = getattr(inner, '_b200_native', None)
            if hit is not None and hit[0] == fp:
                return hit[1]
            nat = AMEDPredictor.from_reference(inner)
            try:
                object.__setattr__(inner, '_b200_native', (fp, nat))
            except Exception:
                pass
            return nat
        return predictor

    def denoise_tap(self, x, i):
        """First evaluation of a step: D(x, t_i) and the channel-mean of the U-Net bottleneck [B, 8, 8]."""
        if self.native:
            D = self.net(x, self.t_dev[i], class_labels=self.kw['class_labels'], out=self.D, bottleneck=self.bott)
            return D, self.bott.reshape(self.B, 8, 8)
        if self.native_ldm:       # middle_block read-out, conditional half (solvers_amed.py:12,24-25)
            D = self.net(x, self.t_dev[i], condition=self.kw['condition'], unconditional_condition=self.kw['unconditional_condition'],
                         out=self.D, bottleneck=self.bott)
            return D, self.bott.reshape(self.B, 8, 8)
        # foreign torch net: fall back to the reference's forward hook (solvers_amed.py:7-18)
        feats = []
        net = self.net
        if hasattr(net, 'guidance_type'):
            mod = net.model.model.diffusion_model.middle_block
        elif getattr(net, 'img_resolution', 0) == 256:
            mod = net.model.middle_block
        else:
            mod = net.model.enc['8x8_block2' if self.kw['class_labels'] is not None else '8x8_block3']
        h = mod.register_forward_hook(lambda m, i_, o: feats.append(o.detach()))
        try:
            D = self.denoise(x, i)
        finally:
            h.remove()
        enc = torch.mean(feats[-1], dim=1)
        if hasattr(net, 'guidance_type') and net.guidance_type == 'classifier-free':
            enc = enc[self.B:]
        return D, enc

    def predict(self, i, enc, use_afs):
        """(r, scale_dir, scale_time) as [B] device vectors (solvers_amed.py:22-55)."""
        from .amed_predictor import AMEDPredictor
        if isinstance(self.pred, AMEDPredictor) and self.latents.device.type == 'cuda':
            o = self.pred.predict_native(None if use_afs else enc, self.t_dev[i], self.t_dev[i + 1], self.B)
            return o[0], o[1], o[2], o[3]
        if use_afs:
            enc = torch.zeros(self.B, 8, 8, device=self.latents.device)
        t_cur = self.t_dev[i].reshape(-1, 1, 1, 1)
        t_next = self.t_dev[i + 1].reshape(-1, 1, 1, 1)
        out = self.pred(enc, t_cur, t_next)
        ones = torch.ones(self.B, device=self.latents.device)
        flat = lambda v: v.reshape(-1).to(torch.float32).expand(self.B) if v.numel() == 1 else v.reshape(-1).to(torch.float32)
        if isinstance(out, (tuple, list)):
            if len(out) == 3:
                r, sd, st = (flat(v) for v in out)
            else:
                owner = getattr(self.pred, 'module', self.pred)
                if owner.scale_time:
                    r, st = flat(out[0]), flat(out[1])
                    sd = ones
                else:
                    r, sd = flat(out[0]), flat(out[1])
                    st = ones
        else:
            r, sd, st = flat(out), ones, ones
        t_mid = (self.t_dev[i + 1] ** r) * (self.t_dev[i] ** (1 - r))
        return r, sd, st, t_mid

    def upd(self, out, xb, coefs, **kw):
        """Fused update with per-sample coefficient vectors (floats are broadcast)."""
        dev = self.latents.device
        rows = [c.to(torch.float32).expand(self.B) if torch.is_tensor(c) else torch.full((self.B,), float(c), device=dev) for c in coefs]
        rows += [torch.zeros(self.B, device=dev)] * (6 - len(rows))
        cd = torch.stack(rows).contiguous()
        return solver_update(out, xb, [0.0] * 6, coef_dev=cd, **kw)


def _mk(predictor, net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max, schedule_type,
        schedule_rho, return_inters, denoise_to_zero):
    return _Amed(predictor, net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
                 schedule_type, schedule_rho, return_inters, denoise_to_zero)


# ---------------------------------------------------------------------------------------------------------------------

@torch.no_grad()
def amed_sampler(net, latents, class_labels=None, condition=None, unconditional_condition=None, num_steps=None, sigma_min=0.002,
                 sigma_max=80, schedule_type='polynomial', schedule_rho=7, afs=False, denoise_to_zero=False, return_inters=False,
                 AMED_predictor=None, step_idx=None, train=False, **kwargs):
    """AMED-Solver: a learned-midpoint single-step method.  Reference: solvers_amed.py:69-159."""
    assert AMED_predictor is not None
    _no_train(train)
    L = _mk(AMED_predictor, net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
            schedule_type, schedule_rho, return_inters, denoise_to_zero)
    t = L.t
    x_mid = torch.empty_like(L.latents)
    for i in range(L.n - 1):
        use_afs = afs and i == 0
        enc = None
        if not use_afs:
            D, enc = L.denoise_tap(L.x, i)
        r, sd, st, t_mid = L.predict(i, enc, use_afs)
        if use_afs:
            L.upd(x_mid, L.x, [1.0, t_mid - t[i]], mode=S.DS_M_DIV, t=_afs_div(t[i]))
        else:
            L.upd(x_mid, L.x, [1.0, t_mid - t[i]], mode=S.DS_M_EPS, D=D, t=t[i])
        D2 = L.denoise(x_mid, sigma=(st * t_mid).contiguous())
        out = L.next_slot(i)
        L.upd(out, L.x, [1.0, sd * (t[i + 1] - t[i])], mode=S.DS_M_EPS, D=D2, xs=x_mid, t_dev=t_mid.contiguous())
        L.x = out
    return L.finish()


@torch.no_grad()
def euler_sampler(net, latents, class_labels=None, condition=None, unconditional_condition=None, num_steps=None, sigma_min=0.002,
                  sigma_max=80, schedule_type='polynomial', schedule_rho=7, afs=False, denoise_to_zero=False, return_inters=False,
                  AMED_predictor=None, step_idx=None, train=False, **kwargs):
    """AMED plug-in for Euler (two Euler legs through the learned intermediate time).  Reference: solvers_amed.py:163-257."""
    _no_train(train)
    if AMED_predictor is None:
        return base.euler_sampler(net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
                                  schedule_type, schedule_rho, afs, denoise_to_zero, return_inters)
    L = _mk(AMED_predictor, net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
            schedule_type, schedule_rho, return_inters, denoise_to_zero)
    t = L.t
    x_mid = torch.empty_like(L.latents)
    for i in range(L.n - 1):
        use_afs = afs and i == 0
        enc = None
        if not use_afs:
            D, enc = L.denoise_tap(L.x, i)
        r, sd, st, t_mid = L.predict(i, enc, use_afs)
        if use_afs:
            L.upd(x_mid, L.x, [1.0, t_mid - t[i]], mode=S.DS_M_DIV, t=_afs_div(t[i]))
        else:
            L.upd(x_mid, L.x, [1.0, t_mid - t[i]], mode=S.DS_M_EPS, D=D, t=t[i])
        D2 = L.denoise(x_mid, sigma=(st * t_mid).contiguous())
        out = L.next_slot(i)
        L.upd(out, x_mid, [1.0, sd * (t[i + 1] - t_mid)], mode=S.DS_M_EPS, D=D2, t_dev=t_mid.contiguous())
        L.x = out
    return L.finish()


@torch.no_grad()
def ipndm_sampler(net, latents, class_labels=None, condition=None, unconditional_condition=None, num_steps=None, sigma_min=0.002,
                  sigma_max=80, schedule_type='polynomial', schedule_rho=7, afs=False, denoise_to_zero=False, return_inters=False,
                  AMED_predictor=None, train=False, max_order=4, buffer_model=[], **kwargs):
    """AMED plug-in for iPNDM.  Reference: solvers_amed.py:262-396.  Both legs use the Adams-Bashforth weights of the
    current history length; the intermediate d enters the history as well."""
    assert max_order >= 1 and max_order <= 4
    _no_train(train)
    if AMED_predictor is None:
        return base.ipndm_sampler(net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
                                  schedule_type, schedule_rho, afs, denoise_to_zero, return_inters, False, max_order)
    L = _mk(AMED_predictor, net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
            schedule_type, schedule_rho, return_inters, denoise_to_zero)
    t = L.t
    hist = []                                    # most recent first, at most max_order - 1 entries
    pool = [torch.empty_like(L.latents) for _ in range(max_order + 1)]
    used = 0
    x_mid = torch.empty_like(L.latents)

    def push(d):
        nonlocal hist
        if max_order > 1:
            hist = ([d] + hist)[:max_order - 1]

    for i in range(L.n - 1):
        use_afs = afs and len(hist) == 0
        enc = None
        if not use_afs:
            D, enc = L.denoise_tap(L.x, i)
        r, sd, st, t_mid = L.predict(i, enc, use_afs)
        order = min(max_order, len(hist) + 1)
        h1 = t_mid - t[i]
        d_cur = pool[used % len(pool)]
        used += 1
        cf = [1.0] + [h1 * a for a in AB[order]]
        if use_afs:
            L.upd(x_mid, L.x, cf, mode=S.DS_M_DIV, t=_afs_div(t[i]), hist=hist[:order - 1], out_m=d_cur)
        else:
            L.upd(x_mid, L.x, cf, mode=S.DS_M_EPS, D=D, t=t[i], hist=hist[:order - 1], out_m=d_cur)
        push(d_cur)
        order = min(max_order, len(hist) + 1)
        D2 = L.denoise(x_mid, sigma=(st * t_mid).contiguous())
        d_mid = pool[used % len(pool)]
        used += 1
        h2 = sd * (t[i + 1] - t_mid)
        out = L.next_slot(i)
        L.upd(out, x_mid, [1.0] + [h2 * a for a in AB[order]], mode=S.DS_M_EPS, D=D2, t_dev=t_mid.contiguous(), hist=hist[:order - 1],
              out_m=d_mid)
        push(d_mid)
        L.x = out
    return L.finish()


@torch.no_grad()
def dpm_2_sampler(net, latents, class_labels=None, condition=None, unconditional_condition=None, num_steps=None, sigma_min=0.002,
                  sigma_max=80, schedule_type='polynomial', schedule_rho=7, afs=False, denoise_to_zero=False, return_inters=False,
                  AMED_predictor=None, step_idx=None, train=False, r=0.5, **kwargs):
    """AMED plug-in for DPM-Solver-2 (learned r).  Reference: solvers_amed.py:400-494."""
    _no_train(train)
    if AMED_predictor is None:
        return base.dpm_2_sampler(net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
                                  schedule_type, schedule_rho, afs, denoise_to_zero, return_inters, False, r)
    L = _mk(AMED_predictor, net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
            schedule_type, schedule_rho, return_inters, denoise_to_zero)
    t = L.t
    x_mid = torch.empty_like(L.latents)
    d_cur = torch.empty_like(L.latents)
    for i in range(L.n - 1):
        use_afs = afs and i == 0
        enc = None
        if not use_afs:
            D, enc = L.denoise_tap(L.x, i)
        rr, sd, st, t_mid = L.predict(i, enc, use_afs)
        if use_afs:
            L.upd(x_mid, L.x, [1.0, t_mid - t[i]], mode=S.DS_M_DIV, t=_afs_div(t[i]), out_m=d_cur)
        else:
            L.upd(x_mid, L.x, [1.0, t_mid - t[i]], mode=S.DS_M_EPS, D=D, t=t[i], out_m=d_cur)
        D2 = L.denoise(x_mid, sigma=(st * t_mid).contiguous())
        h = sd * (t[i + 1] - t[i])
        out = L.next_slot(i)
        L.upd(out, L.x, [1.0, h * (1 / (2 * rr)), h * (1 - 1 / (2 * rr))], mode=S.DS_M_EPS, D=D2, xs=x_mid, t_dev=t_mid.contiguous(),
              hist=[d_cur])
        L.x = out
    return L.finish()


@torch.no_grad()
def dpm_pp_sampler(net, latents, class_labels=None, condition=None, unconditional_condition=None, num_steps=None, sigma_min=0.002,
                   sigma_max=80, schedule_type='polynomial', schedule_rho=7, afs=False, denoise_to_zero=False, return_inters=False,
                   AMED_predictor=None, step_idx=None, train=False, buffer_model=[], buffer_t=[], max_order=3, predict_x0=True,
                   lower_order_final=True, **kwargs):
    """AMED plug-in for multistep DPM-Solver++.  Reference: solvers_amed.py:498-631 with amed-solver-main/solver_utils.py:90-160
    (`scale=` on the second leg).  Every (x, t) pair becomes two multistep updates: t_cur -> t_mid -> t_next."""
    assert max_order >= 1 and max_order <= 3
    _no_train(train)
    if AMED_predictor is None:
        return base.dpm_pp_sampler(net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
                                   schedule_type, schedule_rho, afs, denoise_to_zero, return_inters, False, max_order, predict_x0,
                                   lower_order_final)
    L = _mk(AMED_predictor, net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
            schedule_type, schedule_rho, return_inters, denoise_to_zero)
    t = L.t
    dev = L.latents.device
    B = L.B
    total = 2 * num_steps - 1                                   # solvers_amed.py:560
    hist, hist_t = [], []                                       # most recent first; times as [B] device vectors
    pool = [torch.empty_like(L.latents) for _ in range(5)]
    used = 0
    thr = torch.empty(B, device=dev)
    x_mid = torch.empty_like(L.latents)
    vec = lambda v: torch.full((B,), float(v), device=dev)

    def order_at(step_cur):
        if lower_order_final:
            return step_cur if step_cur < max_order else min(max_order, total - step_cur)
        return min(max_order, step_cur)

    for i in range(L.n - 1):
        step_cur = 2 * i + 1
        use_afs = afs and len(hist) == 0
        enc = None
        m_new = pool[used % 5]
        used += 1
        if use_afs:
            if predict_x0:        # denoised = x - t*d, d = x/sqrt(1+t^2)
                solver_update(L.D, L.x, [1.0, -t[i]], mode=S.DS_M_DIV, t=_afs_div(t[i]))
                D = L.D
        else:
            D, enc = L.denoise_tap(L.x, i)
        r, sd, st, t_mid = L.predict(i, enc, use_afs)
        ts = (hist_t[::-1] + [vec(t[i])])[-3:]
        order = order_at(step_cur)
        c = dpm_pp_coefs(ts, t_mid, order, predict_x0, 1.0, xp=torch)
        cf = [c[0], c[1]] + list(c[2:1 + order])
        if predict_x0:
            dyn_threshold(D, out=thr)
            L.upd(x_mid, L.x, cf, mode=S.DS_M_X0, D=D, thr=thr, hist=hist[:order - 1], out_m=m_new)
        elif use_afs:
            L.upd(x_mid, L.x, cf, mode=S.DS_M_DIV, t=_afs_div(t[i]), hist=hist[:order - 1], out_m=m_new)
        else:
            L.upd(x_mid, L.x, cf, mode=S.DS_M_EPS, D=D, t=t[i], hist=hist[:order - 1], out_m=m_new)
        hist, hist_t = ([m_new] + hist)[:3], ([vec(t[i])] + hist_t)[:3]
        # second leg: evaluate at scale_time * t_mid, step t_mid -> t_next scaled by scale_dir
        step_cur += 1
        D2 = L.denoise(x_mid, sigma=(st * t_mid).contiguous())
        m2 = pool[used % 5]
        used += 1
        ts = (hist_t[::-1] + [t_mid])[-3:]
        order = order_at(step_cur)
        c = dpm_pp_coefs(ts, vec(t[i + 1]), order, predict_x0, sd, xp=torch)
        cf = [c[0], c[1]] + list(c[2:1 + order])
        out = L.next_slot(i)
        if predict_x0:
            dyn_threshold(D2, out=thr)
            L.upd(out, x_mid, cf, mode=S.DS_M_X0, D=D2, thr=thr, hist=hist[:order - 1], out_m=m2)
        else:
            L.upd(out, x_mid, cf, mode=S.DS_M_EPS, D=D2, t_dev=t_mid.contiguous(), hist=hist[:order - 1], out_m=m2)
        hist, hist_t = ([m2] + hist)[:3], ([t_mid] + hist_t)[:3]
        L.x = out
    return L.finish()


@torch.no_grad()
def heun_sampler(net, latents, class_labels=None, condition=None, unconditional_condition=None, num_steps=None, sigma_min=0.002,
                 sigma_max=80, schedule_type='polynomial', schedule_rho=7, afs=False, denoise_to_zero=False, return_inters=False,
                 **kwargs):
    """Plain Heun (teacher sampler).  Reference: solvers_amed.py:635-708."""
    return base.heun_sampler(net, latents, class_labels, condition, unconditional_condition, num_steps, sigma_min, sigma_max,
                             schedule_type, schedule_rho, afs, denoise_to_zero, return_inters)

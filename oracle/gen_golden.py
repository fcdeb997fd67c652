"""Generate tests/golden/*.npz from the REAL reference (/root/reference), run on CPU fp32 in this container.

    python oracle/gen_golden.py            # writes tests/golden/ref_core.npz, ref_amed.npz

The fixtures pin oracle/ (see tests/test_oracle_golden.py); they are small (seeded inputs, tiny nets, a few full-size
outputs).  /root/reference is read-only and never copied: this script only imports it and records its outputs.
Each reference sub-project has its own `solver_utils`, so the AMED part runs in a child interpreter.
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference'

SAMPLER_CASES = [
    ('euler', dict(num_steps=6)),
    ('euler', dict(num_steps=5, afs=True, denoise_to_zero=True)),
    ('heun', dict(num_steps=5)),
    ('dpm_2', dict(num_steps=5, r=0.4)),
    ('ipndm', dict(num_steps=7, max_order=4)),
    ('ipndm', dict(num_steps=6, max_order=3, afs=True)),
    ('ipndm_v', dict(num_steps=7, max_order=4)),
    ('deis', dict(num_steps=7, max_order=4, deis_mode='tab')),
    ('deis', dict(num_steps=6, max_order=4, deis_mode='rhoab')),
    ('dpm_pp', dict(num_steps=7, max_order=3, predict_x0=True)),
    ('dpm_pp', dict(num_steps=6, max_order=2, predict_x0=False)),
    ('dpm_pp', dict(num_steps=6, max_order=3, predict_x0=True, afs=True, lower_order_final=False)),
    ('unipc', dict(num_steps=7, max_order=3, predict_x0=True, variant='bh2')),
    ('unipc', dict(num_steps=6, max_order=2, predict_x0=False, variant='bh1')),
    ('unipc', dict(num_steps=6, max_order=3, predict_x0=True, afs=True)),
]


def digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().to(torch.float32).contiguous().numpy().tobytes())
    return h.hexdigest()


def build_ref_net(name, seed=0, dezero=False):
    """Construct the reference EDMPrecond exactly as sfd-main/training/training_loop.py:62-76 does, random init."""
    sys.path.insert(0, ROOT)
    from oracle import edm_oracle as O
    from models.networks_edm import EDMPrecond
    cfg = dict(O.CONFIGS[name])
    fn = cfg.pop('fn')
    kw = dict(img_resolution=cfg.pop('img_resolution'), img_channels=cfg.pop('in_channels'), label_dim=cfg.pop('label_dim', 0))
    cfg.pop('out_channels')
    cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}
    torch.manual_seed(seed)
    if fn == 'song':
        net = EDMPrecond(model_type='SongUNet', embedding_type='positional', encoder_type='standard', decoder_type='standard',
                         channel_mult_noise=1, resample_filter=[1, 1], dropout=0.13, use_fp16=False, **cfg, **kw)
    else:
        net = EDMPrecond(model_type='DhariwalUNet', **cfg, **kw)
    net.eval().requires_grad_(False)
    if dezero:
        P, _ = O.make_net(name, seed=seed, dezero=True)       # only used as the recipe for the de-zeroed weights
        missing = net.load_state_dict(P, strict=False)
        assert all('resample_filter' in k for k in missing.missing_keys), missing
    return net


def core():
    sys.path.insert(0, os.path.join(REF, 'diff-solvers-main'))
    import solver_utils as RU
    import solvers as RS
    sys.path.insert(0, ROOT)
    from oracle import edm_oracle as O
    G = {}
    meta = {}
    # ---- schedules (solver_utils.py:6-52)
    for st in ('polynomial', 'logsnr', 'time_uniform'):
        for n in (4, 6, 7, 11, 18, 61):
            G[f'sched/{st}/{n}'] = RU.get_schedule(n, 0.002, 80, device=torch.device('cpu'), schedule_type=st, schedule_rho=7).float().numpy()
    # ---- networks: parameter digests and denoiser outputs
    for name in ('tiny_song', 'tiny_adm', 'cifar10'):
        for dz in (False, True):
            net = build_ref_net(name, dezero=dz)
            sd = {k: v for k, v in net.state_dict().items() if 'resample_filter' not in k}
            meta[f'digest/{name}/{int(dz)}'] = digest(sd)
            B = 2
            x = O.stacked_randn(range(B), (3, net.img_resolution, net.img_resolution))
            lab = None
            if net.label_dim:
                lab = torch.eye(net.label_dim)[torch.tensor([1, 3])]
            for sigma in (40.0, 1.0):
                with torch.no_grad():
                    D = net(x * sigma, torch.tensor(sigma), class_labels=lab)
                G[f'net/{name}/{int(dz)}/{sigma}'] = D.numpy()
            sig = torch.tensor([3.0, 0.4])
            with torch.no_grad():
                G[f'net/{name}/{int(dz)}/persample'] = net(x * sig[:, None, None, None], sig, class_labels=lab).numpy()
    # ---- samplers on the de-zeroed tiny nets
    for name in ('tiny_song', 'tiny_adm'):
        net = build_ref_net(name, dezero=True)
        B = 4
        lat = O.stacked_randn(range(B), (3, net.img_resolution, net.img_resolution))
        lab = None
        if net.label_dim:
            g = torch.Generator().manual_seed(0)
            lab = torch.eye(net.label_dim)[torch.randint(net.label_dim, (B,), generator=g)]
        for ci, (solver, kw) in enumerate(SAMPLER_CASES):
            kw = dict(kw)
            mode = kw.pop('deis_mode', None)
            if solver == 'deis':
                ts = RU.get_schedule(kw['num_steps'], 0.002, 80, device=torch.device('cpu'))
                kw['coeff_list'] = RU.get_deis_coeff_list(ts, kw['max_order'], deis_mode=mode)
                G[f'deis/{ci}'] = np.array([[float(c.detach() if torch.is_tensor(c) else c) for c in row] + [0.0] * (4 - len(row)) for row in kw['coeff_list']], dtype=np.float64)
            fn = getattr(RS, solver + '_sampler')
            out = fn(net, lat, class_labels=lab, sigma_min=0.002, sigma_max=80, schedule_type='polynomial', schedule_rho=7, **kw)
            G[f'sample/{name}/{ci}'] = out.numpy()
        rt, re = RS.euler_sampler(net, lat, class_labels=lab, num_steps=6, return_inters=True, return_eps=True, denoise_to_zero=True)
        G[f'traj/{name}/x'], G[f'traj/{name}/eps'] = rt.numpy(), re.numpy()
    # ---- solver math on random tensors
    torch.manual_seed(123)
    x0 = torch.randn(5, 3, 16, 16) * torch.tensor([0.2, 0.8, 1.5, 4.0, 20.0])[:, None, None, None]
    G['thr/in'] = x0.numpy()
    G['thr/out'] = RU.dynamic_thresholding_fn(x0).numpy()
    x = torch.randn(2, 3, 8, 8)
    ms = [torch.randn(2, 3, 8, 8) for _ in range(3)]
    ts = [torch.tensor(5.0), torch.tensor(3.0), torch.tensor(2.0)]
    G['upd/x'] = x.numpy()
    G['upd/ms'] = torch.stack(ms).numpy()
    for order in (1, 2, 3):
        for px0 in (0, 1):
            G[f'upd/dpmpp/{order}/{px0}'] = RU.dpm_pp_update(x, ms, ts, torch.tensor(1.2), order, predict_x0=bool(px0)).numpy()
    # ---- GITS dp (gits-main/gits_utils.py:185-203) on a seeded cost matrix
    sys.path.insert(0, os.path.join(REF, 'gits-main'))
    rng = np.random.RandomState(7)
    cm = np.triu(rng.rand(21, 21).astype(np.float32) * np.linspace(1, 3, 21)[None, :].astype(np.float32), 1)
    G['gits/cost'] = cm
    return G, meta


def gits_part(G, meta):
    """dp() and cal_deviation from gits-main/gits_utils.py (imported with its own solver_utils on the path)."""
    code = r'''
import sys, json, numpy as np, torch
sys.path.insert(0, "/root/reference/gits-main")
import types
sys.modules.setdefault("torch_utils.download_util", types.ModuleType("torch_utils.download_util"))
sys.modules["torch_utils.download_util"].check_file_by_key = lambda *a, **k: None
import gits_utils as GU
d = np.load(sys.argv[1])
cm = d["cost"]
out = {}
for ns, coeff in ((5, 1.0), (7, 1.15), (11, 1.1), (4, 0.9)):
    out[f"{ns}/{coeff}"] = GU.dp(cm, ns, cm.shape[0], coeff)
torch.manual_seed(5)
traj = torch.randn(9, 3, 3, 8, 8).cumsum(0)
dev = GU.cal_deviation(traj, 3, 8, bs=3)
np.savez(sys.argv[2], traj=traj.numpy(), dev=dev.numpy())
print(json.dumps(out))
'''
    tmp_in, tmp_out = '/tmp/_gits_in.npz', '/tmp/_gits_out.npz'
    np.savez(tmp_in, cost=G['gits/cost'])
    r = subprocess.run([sys.executable, '-c', code, tmp_in, tmp_out], capture_output=True, text=True, check=True)
    meta['gits/dp'] = json.loads(r.stdout.strip().splitlines()[-1])
    d = np.load(tmp_out)
    G['gits/traj'], G['gits/dev'] = d['traj'], d['dev']


def amed_part():
    """AMED samplers (amed-solver-main/solvers_amed.py) with a random-init AMED_predictor, in a child interpreter."""
    code = r'''
import sys, numpy as np, torch
sys.path.insert(0, "/root/reference/amed-solver-main")
sys.path.insert(0, "%s")
import solvers_amed as RA
from training.networks import AMED_predictor
from oracle import edm_oracle as O
sys.argv = sys.argv
from oracle.gen_golden import build_ref_net
G = {}
for name in ("tiny_song4",):
    net = build_ref_net(name, dezero=True)
    lat = O.stacked_randn(range(3), (3, net.img_resolution, net.img_resolution))
    for ci, (fn, kw, pk) in enumerate([
        ("amed_sampler", dict(num_steps=4), dict(sampler_stu="amed", scale_dir=0.01, scale_time=0.2)),
        ("euler_sampler", dict(num_steps=4, afs=True), dict(sampler_stu="euler", scale_dir=0.01, scale_time=0.2)),
        ("ipndm_sampler", dict(num_steps=5, max_order=3), dict(sampler_stu="ipndm", scale_dir=0.01, scale_time=0.2)),
        ("dpm_2_sampler", dict(num_steps=4), dict(sampler_stu="dpm", scale_dir=0.0, scale_time=0.2)),
        ("dpm_pp_sampler", dict(num_steps=4, max_order=2, predict_x0=False, afs=True), dict(sampler_stu="dpmpp", scale_dir=0.01, scale_time=0.2)),
        ("dpm_pp_sampler", dict(num_steps=5, max_order=3, predict_x0=True), dict(sampler_stu="dpmpp", scale_dir=0.05, scale_time=0.0)),
    ]):
        torch.manual_seed(100 + ci)
        pred = AMED_predictor(sampler_tea="heun", num_steps=kw["num_steps"], **pk).eval().requires_grad_(False)
        for k, v in pred.state_dict().items():
            G[f"amed/{ci}/pred/{k}"] = v.numpy()
        with torch.no_grad():
            out = getattr(RA, fn)(net, lat, AMED_predictor=pred, sigma_min=0.002, sigma_max=80, schedule_type="polynomial", schedule_rho=7, **kw)
        G[f"amed/{ci}/out"] = out.numpy()
np.savez_compressed(sys.argv[1], **G)
''' % ROOT
    out = os.path.join(OUT, 'ref_amed.npz')
    env = dict(os.environ)
    r = subprocess.run([sys.executable, '-c', code, out], capture_output=True, text=True, env=env)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-4000:])
        raise SystemExit('AMED golden generation failed')
    return out


def amed_taps_part():
    """AMED samplers on the two bottleneck taps ref_amed.npz does not cover (amed-solver-main/solvers_amed.py:11-16, :24-26):
      * class-conditional EDM net -> hook on enc['8x8_block2'] (tiny_adm3, labels given);
      * latent-diffusion net under classifier-free guidance -> hook on model.model.diffusion_model.middle_block, channel mean,
        conditional half [B:] (tiny_ldm, CFG 7.5) with the Stable-Diffusion launch flags (launch.sh:57-61).
    Child interpreter with amed-solver-main on the path; writes tests/golden/ref_amed_taps.npz."""
    code = r"""
import sys, types, numpy as np, torch
sys.path.insert(0, "/root/reference/amed-solver-main")
sys.path.insert(0, "%s")
oc = types.ModuleType("omegaconf"); lc = types.ModuleType("omegaconf.listconfig")
lc.ListConfig = type("ListConfig", (list,), {}); oc.listconfig = lc
sys.modules.setdefault("omegaconf", oc); sys.modules.setdefault("omegaconf.listconfig", lc)
import solvers_amed as RA
import solver_utils as RU
from training.networks import AMED_predictor
from oracle import edm_oracle as O
from oracle import ldm_oracle as LO
from oracle.gen_golden import build_ref_net
G = {}
# ---- class-conditional EDM net: 8x8_block2 tap
net = build_ref_net("tiny_adm3", dezero=True)
B = 3
lat = O.stacked_randn(range(B), (3, net.img_resolution, net.img_resolution))
lab = torch.eye(net.label_dim)[torch.tensor([1, 7, 4])]
G["adm/labels"] = lab.numpy()
for ci, (fn, kw, pk) in enumerate([
    ("amed_sampler", dict(num_steps=4), dict(sampler_stu="amed", scale_dir=0.01, scale_time=0.2)),
    ("dpm_pp_sampler", dict(num_steps=4, max_order=2, predict_x0=True, afs=True), dict(sampler_stu="dpmpp", scale_dir=0.01, scale_time=0.2)),
    ("ipndm_sampler", dict(num_steps=4, max_order=4, afs=True), dict(sampler_stu="ipndm", scale_dir=0.01, scale_time=0.0)),
]):
    torch.manual_seed(200 + ci)
    pred = AMED_predictor(sampler_tea="heun", num_steps=kw["num_steps"], **pk).eval().requires_grad_(False)
    for k, v in pred.state_dict().items():
        G[f"adm/{ci}/pred/{k}"] = v.numpy()
    with torch.no_grad():
        out = getattr(RA, fn)(net, lat, class_labels=lab, AMED_predictor=pred, sigma_min=0.002, sigma_max=80, schedule_type="polynomial",
                              schedule_rho=7, **kw)
    G[f"adm/{ci}/out"] = out.numpy()
# ---- latent-diffusion net, classifier-free guidance: middle_block tap, conditional half
from models.ldm.modules.diffusionmodules.openaimodel import UNetModel
from models.networks_edm import CFGPrecond
P, cfg = LO.make_params("tiny_ldm", seed=0)
unet = UNetModel(image_size=32, in_channels=cfg["in_channels"], out_channels=cfg["out_channels"], model_channels=cfg["model_channels"],
                 attention_resolutions=list(cfg["attention_resolutions"]), num_res_blocks=cfg["num_res_blocks"],
                 channel_mult=list(cfg["channel_mult"]), num_heads=cfg["num_heads"], use_spatial_transformer=True,
                 transformer_depth=1, context_dim=cfg["context_dim"], use_checkpoint=False, legacy=False).eval().requires_grad_(False)
unet.load_state_dict(P)
class Inner(torch.nn.Module):
    def __init__(self, u):
        super().__init__()
        self.diffusion_model = u
class Shim(torch.nn.Module):
    def __init__(self, u):
        super().__init__()
        self.model = Inner(u)
        self.alphas_cumprod = LO.make_alphas_cumprod()
    def apply_model(self, x, t, cond):
        return self.model.diffusion_model(x, t, context=cond)
R = cfg["img_resolution"]
B = 2
lat = O.stacked_randn(range(B), (cfg["in_channels"], R, R))
g = torch.Generator().manual_seed(9)
c = torch.randn(B, 77, cfg["context_dim"], generator=g)
uc = torch.randn(B, 77, cfg["context_dim"], generator=g)
G["ldm/c"], G["ldm/uc"] = c.numpy(), uc.numpy()
for ci, (guidance, kw, pk) in enumerate([
    (7.5, dict(num_steps=4, afs=True, max_order=2, predict_x0=False, lower_order_final=True), dict(sampler_stu="dpmpp", scale_dir=0.0, scale_time=0.2)),
    (7.5, dict(num_steps=3, afs=False, max_order=3, predict_x0=False, lower_order_final=True), dict(sampler_stu="dpmpp", scale_dir=0.01, scale_time=0.2)),
    # (guidance 1.0 cannot be recorded: the reference multiplies c_in[B] * x without a reshape on that branch, networks_edm.py:685,
    #  and the AMED second evaluation always carries a per-sample sigma)
]):
    net = CFGPrecond(Shim(unet), img_resolution=R, img_channels=cfg["in_channels"], guidance_rate=guidance,
                     guidance_type="classifier-free", label_dim=True).eval()
    torch.manual_seed(300 + ci)
    pred = AMED_predictor(sampler_tea="dpmpp", num_steps=kw["num_steps"], **pk).eval().requires_grad_(False)
    for k, v in pred.state_dict().items():
        G[f"ldm/{ci}/pred/{k}"] = v.numpy()
    with torch.no_grad():
        out = RA.dpm_pp_sampler(net, lat, condition=c, unconditional_condition=uc, AMED_predictor=pred, sigma_min=net.sigma_min,
                                sigma_max=net.sigma_max, schedule_type="discrete", schedule_rho=1, **kw)
    G[f"ldm/{ci}/out"] = out.numpy()
np.savez_compressed(sys.argv[1], **G)
""" % ROOT
    out = os.path.join(OUT, 'ref_amed_taps.npz')
    r = subprocess.run([sys.executable, '-c', code, out], capture_output=True, text=True, env=dict(os.environ))
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-4000:])
        raise SystemExit('AMED tap golden generation failed')
    return out


def ldm_part():
    """Stable-Diffusion-style eps-net + CFGPrecond from the real reference, loaded with oracle/ldm_oracle.make_params weights."""
    import types
    sys.path.insert(0, os.path.join(REF, 'diff-solvers-main'))
    sys.path.insert(0, ROOT)
    oc = types.ModuleType('omegaconf')
    lc = types.ModuleType('omegaconf.listconfig')
    lc.ListConfig = type('ListConfig', (list,), {})
    oc.listconfig = lc
    sys.modules.setdefault('omegaconf', oc)
    sys.modules.setdefault('omegaconf.listconfig', lc)
    from models.ldm.modules.diffusionmodules.openaimodel import UNetModel
    from models.networks_edm import CFGPrecond
    from oracle import edm_oracle as O
    from oracle import ldm_oracle as LO
    G = {}
    for name in ('tiny_ldm',):
        P, cfg = LO.make_params(name, seed=0)
        unet = UNetModel(image_size=32, in_channels=cfg['in_channels'], out_channels=cfg['out_channels'], model_channels=cfg['model_channels'],
                         attention_resolutions=list(cfg['attention_resolutions']), num_res_blocks=cfg['num_res_blocks'],
                         channel_mult=list(cfg['channel_mult']), num_heads=cfg['num_heads'], use_spatial_transformer=True,
                         transformer_depth=1, context_dim=cfg['context_dim'], use_checkpoint=False, legacy=False).eval().requires_grad_(False)
        sd = unet.state_dict()
        assert list(sd.keys()) == list(P.keys()), [k for k in sd if k not in P][:5] + [k for k in P if k not in sd][:5]
        for k in sd:
            assert tuple(sd[k].shape) == tuple(P[k].shape), (k, sd[k].shape, P[k].shape)
        unet.load_state_dict(P)

        class Shim(torch.nn.Module):
            def __init__(self, u):
                super().__init__()
                self.u = u
                self.alphas_cumprod = LO.make_alphas_cumprod()

            def apply_model(self, x, t, cond):
                return self.u(x, t, context=cond)
        net = CFGPrecond(Shim(unet), img_resolution=cfg['img_resolution'], img_channels=cfg['in_channels'], guidance_rate=7.5,
                         guidance_type='classifier-free', label_dim=True).eval()
        G[f'ldm/{name}/sigma_range'] = np.array([net.sigma_min, net.sigma_max])
        B = 2
        R = cfg['img_resolution']
        x = O.stacked_randn(range(B), (cfg['in_channels'], R, R))
        g = torch.Generator().manual_seed(5)
        c = torch.randn(B, 77, cfg['context_dim'], generator=g)
        uc = torch.randn(B, 77, cfg['context_dim'], generator=g)
        G[f'ldm/{name}/c'], G[f'ldm/{name}/uc'] = c.numpy(), uc.numpy()
        with torch.no_grad():
            for sigma in (10.0, 0.5):
                G[f'ldm/{name}/D/{sigma}'] = net(x * sigma, torch.tensor([sigma]), condition=c, unconditional_condition=uc).numpy()
            G[f'ldm/{name}/D/nocfg'] = net(x * 2.0, torch.tensor([2.0]), condition=c, unconditional_condition=None).numpy()
            sig = torch.tensor([3.0, 0.4])
            G[f'ldm/{name}/D/persample'] = net(x * sig[:, None, None, None], sig, condition=c, unconditional_condition=uc).numpy()
            G[f'ldm/{name}/eps'] = unet(x, torch.tensor([500.0, 20.0]), context=c).numpy()
            sys.path.insert(0, os.path.join(REF, 'diff-solvers-main'))
            import solver_utils as RU
            import solvers as RS
            ts = RU.get_schedule(5, net.sigma_min, net.sigma_max, device=torch.device('cpu'), schedule_type='discrete', schedule_rho=1, net=net)
            G[f'ldm/{name}/sched_discrete'] = ts.numpy()
            out = RS.dpm_pp_sampler(net, x, condition=c, unconditional_condition=uc, num_steps=5, sigma_min=net.sigma_min, sigma_max=net.sigma_max,
                                    schedule_type='discrete', schedule_rho=1, max_order=2, predict_x0=False)
            G[f'ldm/{name}/sample_dpmpp'] = out.numpy()
    # known-answer anchor from the reference docs (amed-solver-main/example.ipynb): sigma range of SD
    out = os.path.join(OUT, 'ref_ldm.npz')
    np.savez_compressed(out, **G)
    return out


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if len(sys.argv) > 1 and sys.argv[1] == 'amed_taps':      # only the round-2 addition (the other fixtures are unchanged)
        print('wrote', amed_taps_part())
        raise SystemExit(0)
    G, meta = core()
    gits_part(G, meta)
    G['meta_json'] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, 'ref_core.npz'), **G)
    print('wrote ref_core.npz with', len(G), 'arrays')
    print('wrote', amed_part())
    print('wrote', ldm_part())
    print('wrote', amed_taps_part())

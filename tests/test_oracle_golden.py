"""Pin the oracle (oracle/*.py, the CPU restatement) against golden vectors recorded from the REAL reference
(tests/golden/ref_core.npz, produced by oracle/gen_golden.py importing /root/reference).  CPU only.

The reference ships no tests of its own; these fixtures are the reference's outputs on seeded
inputs.  Network outputs and sampler results must match bit-for-bit or to fp32 round-off (the oracle executes the
same torch CPU ops in the same order); integer outputs (GITS dp lists) must be identical.
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import edm_oracle as O
from oracle import solvers_oracle as SO

GOLD = os.path.join(os.path.dirname(__file__), 'golden', 'ref_core.npz')


@pytest.fixture(scope='module')
def gold():
    d = np.load(GOLD)
    meta = json.loads(bytes(d['meta_json']).decode())
    return d, meta


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


def test_schedules_bit_exact(gold):
    d, _ = gold
    for st in ('polynomial', 'logsnr', 'time_uniform'):
        for n in (4, 6, 7, 11, 18, 61):
            ref = d[f'sched/{st}/{n}']
            got = SO.get_schedule(n, 0.002, 80, schedule_type=st, schedule_rho=7).float().numpy()
            assert np.array_equal(ref, got), (st, n)
    # known-answer anchor quoted in SURVEY.md section 8(c)
    ks = SO.get_schedule(6, 0.002, 80).numpy()
    assert np.allclose(ks, [80, 24.4083, 5.83894, 0.965417, 0.0850872, 0.002], rtol=2e-5)


@pytest.mark.parametrize('name', ['tiny_song', 'tiny_adm', 'cifar10'])
@pytest.mark.parametrize('dz', [0, 1])
def test_network_init_and_forward(gold, name, dz):
    d, meta = gold
    P, S = O.make_net(name, seed=0, dezero=bool(dz))
    assert O.params_digest(P) == meta[f'digest/{name}/{dz}'], 'parameter init differs from the reference constructors'
    net = O.OracleNet(P, S)
    x = O.stacked_randn(range(2), (3, S['img_resolution'], S['img_resolution']))
    lab = torch.eye(S['label_dim'])[torch.tensor([1, 3])] if S['label_dim'] else None
    for sigma in (40.0, 1.0):
        got = net(x * sigma, torch.tensor(sigma), class_labels=lab).numpy()
        ref = d[f'net/{name}/{dz}/{sigma}']
        assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max()), (name, dz, sigma, np.abs(got - ref).max())
    sig = torch.tensor([3.0, 0.4])
    got = net(x * sig[:, None, None, None], sig, class_labels=lab).numpy()
    ref = d[f'net/{name}/{dz}/persample']
    assert np.abs(got - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize('name', ['tiny_song', 'tiny_adm'])
def test_samplers_match_reference(gold, name):
    d, _ = gold
    P, S = O.make_net(name, seed=0, dezero=True)
    net = O.OracleNet(P, S)
    B = 4
    lat = O.stacked_randn(range(B), (3, S['img_resolution'], S['img_resolution']))
    lab = None
    if S['label_dim']:
        g = torch.Generator().manual_seed(0)
        lab = torch.eye(S['label_dim'])[torch.randint(S['label_dim'], (B,), generator=g)]
    for ci, (solver, kw) in enumerate(SAMPLER_CASES):
        kw = dict(kw)
        mode = kw.pop('deis_mode', None)
        if solver == 'deis':
            ts = SO.get_schedule(kw['num_steps'], 0.002, 80)
            kw['coeff_list'] = SO.get_deis_coeff_list(ts, kw['max_order'], deis_mode=mode)
            ref_c = d[f'deis/{ci}']
            for row, rr in zip(kw['coeff_list'], ref_c):
                assert np.allclose([float(c) for c in row], rr[:len(row)], rtol=1e-6, atol=1e-9), (ci, row, rr)
        got = SO.sample(net, lat, solver, class_labels=lab, sigma_min=0.002, sigma_max=80, schedule_type='polynomial', schedule_rho=7, **kw)
        ref = d[f'sample/{name}/{ci}']
        err = np.abs(got.numpy() - ref).max()
        assert err <= 2e-5 * max(1.0, np.abs(ref).max()), (name, solver, kw, err)
    rt, re = SO.sample(net, lat, 'euler', class_labels=lab, num_steps=6, return_inters=True, return_eps=True, denoise_to_zero=True)
    assert rt.shape == d[f'traj/{name}/x'].shape and re.shape == d[f'traj/{name}/eps'].shape
    assert np.abs(rt.numpy() - d[f'traj/{name}/x']).max() < 1e-4
    assert np.abs(re.numpy() - d[f'traj/{name}/eps']).max() < 1e-4


def test_solver_math(gold):
    d, _ = gold
    x0 = torch.from_numpy(d['thr/in'])
    assert np.array_equal(SO.dynamic_thresholding(x0).numpy(), d['thr/out'])
    # the sort-based restatement of torch.quantile used to validate the CUDA radix select
    q = SO.quantile_by_sort(x0.abs().reshape(5, -1), 0.995)
    assert torch.equal(q, torch.quantile(x0.abs().reshape(5, -1), 0.995, dim=1))
    x = torch.from_numpy(d['upd/x'])
    ms = list(torch.from_numpy(d['upd/ms']))
    ts = [torch.tensor(5.0), torch.tensor(3.0), torch.tensor(2.0)]
    for order in (1, 2, 3):
        for px0 in (0, 1):
            got = SO.dpm_pp_update(x, ms, ts, torch.tensor(1.2), order, predict_x0=bool(px0)).numpy()
            assert np.abs(got - d[f'upd/dpmpp/{order}/{px0}']).max() < 2e-6


def test_gits_dp_and_deviation_bit_exact(gold):
    d, meta = gold
    cm = d['gits/cost']
    for key, ref in meta['gits/dp'].items():
        ns, coeff = key.split('/')
        assert SO.dp(cm, int(ns), cm.shape[0], float(coeff)) == ref, key
    traj = torch.from_numpy(d['gits/traj'])
    dev = SO.cal_deviation(traj, 3, 8, bs=3).numpy()
    assert np.abs(dev - d['gits/dev']).max() <= 1e-5 * np.abs(d['gits/dev']).max()


AMED_CASES = [
    ('amed', dict(num_steps=4), dict(scale_dir=0.01, scale_time=0.2)),
    ('euler', dict(num_steps=4, afs=True), dict(scale_dir=0.01, scale_time=0.2)),
    ('ipndm', dict(num_steps=5, max_order=3), dict(scale_dir=0.01, scale_time=0.2)),
    ('dpm_2', dict(num_steps=4), dict(scale_dir=0.0, scale_time=0.2)),
    ('dpm_pp', dict(num_steps=4, max_order=2, predict_x0=False, afs=True), dict(scale_dir=0.01, scale_time=0.2)),
    ('dpm_pp', dict(num_steps=5, max_order=3, predict_x0=True), dict(scale_dir=0.05, scale_time=0.0)),
]


def load_amed_case(ci):
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_amed.npz'))
    W = {k[len(f'amed/{ci}/pred/'):]: torch.from_numpy(d[k]) for k in d.files if k.startswith(f'amed/{ci}/pred/')}
    return W, d[f'amed/{ci}/out']


@pytest.mark.parametrize('ci', range(len(AMED_CASES)))
def test_amed_samplers_match_reference(ci):
    from oracle import amed_oracle as AO
    solver, kw, cfg = AMED_CASES[ci]
    W, ref = load_amed_case(ci)
    P, S = O.make_net('tiny_song4', seed=0, dezero=True)
    net = O.OracleNet(P, S)
    lat = O.stacked_randn(range(3), (3, 16, 16))
    got = AO.sample_amed(net, lat, solver, W, cfg, **kw).numpy()
    err = np.abs(got - ref).max()
    assert err <= 2e-5 * max(1.0, np.abs(ref).max()), (solver, kw, err)


# AMED on the two other bottleneck taps (round 2): class-conditional EDM net -> enc['8x8_block2']; latent diffusion under
# classifier-free guidance -> middle_block, conditional half (amed-solver-main/solvers_amed.py:11-16, :24-26).
AMED_ADM_CASES = [
    ('amed', dict(num_steps=4), dict(scale_dir=0.01, scale_time=0.2)),
    ('dpm_pp', dict(num_steps=4, max_order=2, predict_x0=True, afs=True), dict(scale_dir=0.01, scale_time=0.2)),
    ('ipndm', dict(num_steps=4, max_order=4, afs=True), dict(scale_dir=0.01, scale_time=0.0)),
]
AMED_LDM_CASES = [
    (7.5, dict(num_steps=4, afs=True, max_order=2, predict_x0=False, lower_order_final=True), dict(scale_dir=0.0, scale_time=0.2)),   # launch.sh:57-61
    (7.5, dict(num_steps=3, afs=False, max_order=3, predict_x0=False, lower_order_final=True), dict(scale_dir=0.01, scale_time=0.2)),
]


def load_amed_tap_case(group, ci):
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_amed_taps.npz'))
    pre = f'{group}/{ci}/pred/'
    W = {k[len(pre):]: torch.from_numpy(d[k]) for k in d.files if k.startswith(pre)}
    return d, W, d[f'{group}/{ci}/out']


@pytest.mark.parametrize('ci', range(len(AMED_ADM_CASES)))
def test_amed_class_conditional_tap_matches_reference(ci):
    from oracle import amed_oracle as AO
    solver, kw, cfg = AMED_ADM_CASES[ci]
    d, W, ref = load_amed_tap_case('adm', ci)
    P, S = O.make_net('tiny_adm3', seed=0, dezero=True)
    net = O.OracleNet(P, S)
    lat = O.stacked_randn(range(3), (3, 16, 16))
    got = AO.sample_amed(net, lat, solver, W, cfg, class_labels=torch.from_numpy(d['adm/labels']), **kw).numpy()
    err = np.abs(got - ref).max()
    assert err <= 2e-5 * max(1.0, np.abs(ref).max()), (solver, kw, err)


@pytest.mark.parametrize('ci', range(len(AMED_LDM_CASES)))
def test_amed_ldm_cfg_tap_matches_reference(ci):
    from oracle import amed_oracle as AO
    from oracle import ldm_oracle as LO
    guidance, kw, cfg = AMED_LDM_CASES[ci]
    d, W, ref = load_amed_tap_case('ldm', ci)
    P, lcfg = LO.make_params('tiny_ldm')
    net = LO.OracleCFGNet(P, lcfg, guidance_rate=guidance)
    lat = O.stacked_randn(range(2), (4, 16, 16))
    got = AO.sample_amed(net, lat, 'dpm_pp', W, cfg, condition=torch.from_numpy(d['ldm/c']), unconditional_condition=torch.from_numpy(d['ldm/uc']),
                         sigma_min=net.sigma_min, sigma_max=net.sigma_max, schedule_type='discrete', schedule_rho=1, **kw).numpy()
    err = np.abs(got - ref).max()
    assert err <= 2e-5 * max(1.0, np.abs(ref).max()), (kw, err)


def test_ldm_oracle_matches_reference():
    """Stable-Diffusion-style eps-net + CFGPrecond (tiny config, same structure as v1.5) vs outputs of the real reference classes."""
    from oracle import ldm_oracle as LO
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_ldm.npz'))
    P, cfg = LO.make_params('tiny_ldm')
    net = LO.OracleCFGNet(P, cfg)
    assert abs(net.sigma_min - 0.0292) < 1e-4 and abs(net.sigma_max - 14.6146) < 1e-3        # amed-solver-main/example.ipynb
    assert np.allclose([net.sigma_min, net.sigma_max], d['ldm/tiny_ldm/sigma_range'], rtol=1e-6)
    x = O.stacked_randn(range(2), (4, 16, 16))
    c, uc = torch.from_numpy(d['ldm/tiny_ldm/c']), torch.from_numpy(d['ldm/tiny_ldm/uc'])
    with torch.no_grad():
        for sigma in (10.0, 0.5):
            got = net(x * sigma, torch.tensor([sigma]), condition=c, unconditional_condition=uc).numpy()
            ref = d[f'ldm/tiny_ldm/D/{sigma}']
            assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
        got = net(x * 2.0, torch.tensor([2.0]), condition=c).numpy()
        assert np.abs(got - d['ldm/tiny_ldm/D/nocfg']).max() <= 1e-5 * np.abs(d['ldm/tiny_ldm/D/nocfg']).max()
        sig = torch.tensor([3.0, 0.4])
        got = net(x * sig[:, None, None, None], sig, condition=c, unconditional_condition=uc).numpy()
        assert np.abs(got - d['ldm/tiny_ldm/D/persample']).max() <= 1e-5 * np.abs(d['ldm/tiny_ldm/D/persample']).max()
        eps = LO.unet_forward(P, cfg, x, torch.tensor([500.0, 20.0]), c).numpy()
        assert np.abs(eps - d['ldm/tiny_ldm/eps']).max() <= 1e-5 * np.abs(d['ldm/tiny_ldm/eps']).max()
        ts = SO.get_schedule(5, net.sigma_min, net.sigma_max, schedule_type='discrete', schedule_rho=1, net=net)
        assert np.array_equal(ts.numpy(), d['ldm/tiny_ldm/sched_discrete'])
        out = SO.sample(net, x, 'dpm_pp', condition=c, unconditional_condition=uc, num_steps=5, sigma_min=net.sigma_min,
                        sigma_max=net.sigma_max, schedule_type='discrete', schedule_rho=1, max_order=2, predict_x0=False).numpy()
        assert np.abs(out - d['ldm/tiny_ldm/sample_dpmpp']).max() <= 1e-5 * np.abs(d['ldm/tiny_ldm/sample_dpmpp']).max()


def test_vae_oracle_matches_reference():
    """First-stage decoder restatement (oracle/vae_oracle.py) vs the real reference Decoder run in this container (oracle/gen_vae_golden.py)."""
    from oracle import vae_oracle as VO
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_vae.npz'))
    for name in ('tiny_vae', 'wide_vae'):
        P, cfg = VO.make_params(name, seed=0)
        z = torch.from_numpy(d[f'vae/{name}/z'])
        with torch.no_grad():
            x = VO.decode(P, cfg, z).numpy()
        ref = d[f'vae/{name}/x']
        assert x.shape == ref.shape
        assert np.abs(x - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), np.abs(x - ref).max()


def test_clip_oracle_matches_transformers():
    """Text-encoder restatement (oracle/clip_oracle.py) vs Hugging Face transformers' own CLIPTextModel -- the module the reference's
    FrozenCLIPEmbedder wraps -- run in this container (oracle/gen_clip_golden.py)."""
    from oracle import clip_oracle as CO
    d = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_clip.npz'))
    P, cfg = CO.make_params('tiny_clip', seed=0)
    ids = torch.from_numpy(d['tiny_clip/ids'])
    with torch.no_grad():
        got = CO.text_forward(P, cfg, ids).numpy()
    ref = d['tiny_clip/out']
    assert got.shape == ref.shape == (3, 77, cfg['hidden_size'])
    assert np.abs(got - ref).max() < 2e-5
    # causal: changing a later token leaves earlier positions bit-identical
    ids2 = ids.clone()
    ids2[:, 40] = (ids2[:, 40] + 1) % cfg['vocab_size']
    with torch.no_grad():
        got2 = CO.text_forward(P, cfg, ids2).numpy()
    assert np.array_equal(got2[:, :40], got[:, :40]) and np.abs(got2[:, 40:] - got[:, 40:]).max() > 1e-3

"""World-size-2 (gloo, CPU) coverage of the N > 1 path: the seed partition of sample.py:166-169, the end-of-run gather of
finished uint8 images, and the cost-matrix all_reduce GITS performs (gits_utils.py:134).  No collective exists inside the
sampling loop itself."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from diff_sampler_b200 import dist_utils
    r, w = dist_utils.init(backend='gloo')
    assert (r, w) == (rank, world)
    seeds = list(range(37))
    mine = dist_utils.rank_batches(seeds, 8, world, rank)
    # each rank "samples" its batches: image value encodes the seed
    imgs = [torch.full((len(b), 3, 4, 4), 0.0) + (b.float()[:, None, None, None] - 128) / 127.5 for b in mine]
    u8 = [dist_utils.to_uint8_nhwc(i) for i in imgs]
    # equal shapes are required by all_gather: pad to the largest per-rank batch
    padded = torch.zeros(3, 8, 4, 4, 3, dtype=torch.uint8)
    counts = torch.zeros(3, dtype=torch.int64)
    for k, t in enumerate(u8):
        padded[k, :len(t)] = t
        counts[k] = len(t)
    allimg = dist_utils.gather_images(padded.reshape(-1, 4, 4, 3))
    allcnt = [torch.empty_like(counts) for _ in range(world)]
    dist.all_gather(allcnt, counts)
    cost = torch.full((5, 5), float(rank + 1))
    dist.all_reduce(cost)
    if rank == 0:
        got = []
        per = allimg.reshape(world, 3, 8, 4, 4, 3)
        for rr in range(world):
            for k in range(3):
                got += per[rr, k, :int(allcnt[rr][k]), 0, 0, 0].tolist()
        torch.save(dict(seeds=sorted(got), cost=cost, nb=[len(b) for b in mine]), out)
    dist.barrier()
    dist.destroy_process_group()


def test_seed_partition_gather_and_allreduce(tmp_path):
    out = str(tmp_path / 'r0.pt')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out)
    assert r['seeds'] == list(range(37))                 # every seed sampled exactly once across ranks
    assert torch.equal(r['cost'], torch.full((5, 5), 3.0))
    assert sum(r['nb']) > 0


def test_rank_batches_matches_reference_partition():
    from diff_sampler_b200 import dist_utils
    seeds = torch.arange(50000)
    for world in (1, 2, 4, 8):
        nb = ((len(seeds) - 1) // (512 * world) + 1) * world
        ref = seeds.tensor_split(nb)
        seen = []
        for rank in range(world):
            mine = dist_utils.rank_batches(seeds, 512, world, rank)
            assert all(torch.equal(a, b) for a, b in zip(mine, ref[rank::world]))
            seen += [int(v) for b in mine for v in b]
        assert sorted(seen) == list(range(50000))


def _fid_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from diff_sampler_b200 import dist_utils, fid_stats
    dist_utils.init(backend='gloo')
    g = torch.Generator().manual_seed(5)
    images = torch.randint(0, 256, (41, 6, 6, 3), generator=g, dtype=torch.uint8)         # the "finished samples" of the whole job
    proj = torch.randn(3 * 6 * 6, 16, generator=g, dtype=torch.float64)
    detector = lambda x: (x.reshape(x.shape[0], -1).to(torch.float64) / 255.0) @ proj       # stands in for Inception features
    st = fid_stats.FeatureStats()
    for b in dist_utils.rank_batches(range(41), 8, world, rank):                            # this rank's batches, ragged (41 = 5*8 + 1)
        st.append_images(images[b], detector)
    mu, sigma = st.reduce().finalize()
    if rank == 0:
        torch.save(dict(mu=mu, sigma=sigma, n=st.n, images=images, proj=proj), out)
    dist.barrier()
    dist.destroy_process_group()


def test_fid_statistics_allreduce_matches_single_process_reference_formulas(tmp_path):
    """fid.py:61-79 on two ranks == the same formulas over all images in one process; FID of identical statistics is ~0."""
    import numpy as np
    from diff_sampler_b200 import fid_stats
    out = str(tmp_path / 'fid.pt')
    mp.spawn(_fid_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    r = torch.load(out, weights_only=False)
    assert r['n'] == 41
    f = ((r['images'].permute(0, 3, 1, 2).reshape(41, -1).to(torch.float64) / 255.0) @ r['proj']).numpy()
    mu = f.sum(0) / 41
    sigma = (f.T @ f - np.outer(mu, mu) * 41) / 40
    assert np.allclose(r['mu'], mu, rtol=1e-12, atol=1e-12) and np.allclose(r['sigma'], sigma, rtol=1e-10, atol=1e-10)
    assert np.allclose(sigma, np.cov(f, rowvar=False), rtol=1e-9, atol=1e-9)               # it is the unbiased covariance
    assert abs(fid_stats.frechet_distance(r['mu'], r['sigma'], mu, sigma)) < 1e-6
    shifted = fid_stats.frechet_distance(r['mu'] + 0.5, r['sigma'], mu, sigma)
    assert abs(shifted - 0.25 * 16) < 1e-6                                                  # |dmu|^2 with equal covariances

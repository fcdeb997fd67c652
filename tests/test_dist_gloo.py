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

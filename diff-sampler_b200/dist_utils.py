"""Multi-GPU plumbing of the sampling path: the image batch shards embarrassingly — one process per GPU, a full weight
replica each, NO collective inside the sampling loop.  Only the finished samples are gathered (what fid.py consumes).

Reference: diff-solvers-main/sample.py:166-169 (seed partition), :268,:319 (per-batch barrier),
torch_utils/distributed.py:14-31 (env:// NCCL init).  Works with any torch.distributed backend (NCCL on GPUs, gloo in CPU tests).
"""
import os

import torch
import torch.distributed as dist


def init(backend=None, device=None):
    """env:// process-group init with the reference's single-node defaults (distributed.py:14-27)."""
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29500')
    os.environ.setdefault('RANK', '0')
    os.environ.setdefault('LOCAL_RANK', '0')
    os.environ.setdefault('WORLD_SIZE', '1')
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if not dist.is_initialized():
        kw = {}
        if backend == 'nccl' and device is not None:
            kw['device_id'] = device
        dist.init_process_group(backend=backend, init_method='env://', **kw)
    return dist.get_rank(), dist.get_world_size()


def rank_batches(seeds, max_batch_size, world_size, rank):
    """Split the seed list into per-rank batches exactly as the reference does (sample.py:166-169):
    ceil(len/(batch*W))*W batches by tensor_split, rank r takes batches r::W."""
    seeds = torch.as_tensor(list(seeds))
    num_batches = ((len(seeds) - 1) // (max_batch_size * world_size) + 1) * world_size
    all_batches = seeds.tensor_split(num_batches)
    return all_batches[rank::world_size]


def to_uint8_nhwc(images):
    """(x * 127.5 + 128).clip(0, 255).uint8, NCHW -> NHWC (sample.py:311).  CUDA tensors go through the one-pass native kernel
    (ds_images_to_uint8); CPU tensors (tests of the gather logic) use the torch expression."""
    if images.device.type == 'cuda':
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        x = images.to(torch.float32).contiguous()
        B, Cc, H, W = x.shape
        out = torch.empty(B, H, W, Cc, dtype=torch.uint8, device=x.device)
        stream = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(lib.ds_images_to_uint8(x.data_ptr(), out.data_ptr(), B, Cc, H * W, stream), 'ds_images_to_uint8')
        return out
    return (images * 127.5 + 128).clip(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def gather_images(images_u8):
    """all_gather of equally-shaped uint8 image batches -> [world * B, H, W, C] on every rank (rank-major order)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return images_u8
    parts = [torch.empty_like(images_u8) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, images_u8.contiguous())
    return torch.cat(parts, dim=0)

"""Generate tests/golden/edm_snapshot_*.pkl with the REAL reference (/root/reference), in this container:

    python oracle/gen_edm_pickle.py

Small random-init EDMPrecond networks built from the reference classes (models/networks_edm.py, decorated by
torch_utils/persistence.py) and pickled exactly like EDM's training loop writes `network-snapshot-*.pkl`
(`pickle.dump(dict(ema=net, ...))`), plus tests/golden/edm_snapshot.json with the state_dict digest and attributes the
importer (diff-sampler_b200/checkpoint.py) must reproduce.  Real checkpoints are unreachable here (no network), so these
stand in for the file format; the format does not depend on the network size.

persistence.py embeds the defining module's SOURCE TEXT in every pickle (meta['module_src']); the fixtures must not carry a copy of
the reference's networks_edm.py, so the string is replaced by a one-line placeholder before dumping.  The record structure
(reconstruct hook, meta keys, state dicts) is untouched -- and the importer never reads module_src anyway.
"""
import hashlib
import json
import os
import pickle
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden')
sys.path.insert(0, '/root/reference/diff-solvers-main')


def digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().to(torch.float32).contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    import models.networks_edm as ref_nets
    from models.networks_edm import EDMPrecond, SongUNet
    for obj in vars(ref_nets).values():
        if isinstance(obj, type) and hasattr(obj, '_orig_module_src'):
            obj._orig_module_src = '# module source elided from the test fixture (see oracle/gen_edm_pickle.py)\n'
    meta = {}
    cases = {
        'song': dict(cls=EDMPrecond, kw=dict(img_resolution=8, img_channels=3, label_dim=0, model_type='SongUNet', model_channels=8,
                                             channel_mult=[1, 2], num_blocks=1, attn_resolutions=[4], augment_dim=9, embedding_type='positional',
                                             encoder_type='standard', decoder_type='standard', channel_mult_noise=1, resample_filter=[1, 1])),
        'adm_fp16': dict(cls=EDMPrecond, kw=dict(img_resolution=8, img_channels=3, label_dim=10, use_fp16=True, model_type='DhariwalUNet',
                                                 model_channels=8, channel_mult=[1, 2], num_blocks=1, attn_resolutions=[4])),
        # 64-channel net: the smallest the native kernels run (channel counts are multiples of 64), for the GPU from_pickle test
        'song64': dict(cls=EDMPrecond, kw=dict(img_resolution=8, img_channels=3, label_dim=0, model_type='SongUNet', model_channels=64,
                                               channel_mult=[1], num_blocks=1, attn_resolutions=[8], augment_dim=9, embedding_type='positional',
                                               encoder_type='standard', decoder_type='standard', channel_mult_noise=1, resample_filter=[1, 1])),
        # not an EDMPrecond: the importer must refuse it (a bare U-Net has no preconditioning to drop in for)
        'bare_unet': dict(cls=SongUNet, kw=dict(img_resolution=8, in_channels=3, out_channels=3, model_channels=8, channel_mult=[1],
                                                num_blocks=1, attn_resolutions=[], resample_filter=[1, 1])),
    }
    for name, c in cases.items():
        torch.manual_seed(7)
        net = c['cls'](**c['kw']).eval().requires_grad_(False)
        path = os.path.join(OUT, f'edm_snapshot_{name}.pkl')
        with open(path, 'wb') as f:
            pickle.dump(dict(ema=net, loss_fn=None, augment_pipe=None, dataset_kwargs=dict(resolution=8)), f)
        sd = {k: v for k, v in net.state_dict().items() if 'resample_filter' not in k}
        meta[name] = dict(file=os.path.basename(path), class_name=type(net).__name__, digest=digest(sd), n_tensors=len(sd),
                          keys_head=list(sd.keys())[:4], img_resolution=getattr(net, 'img_resolution', None), img_channels=getattr(net, 'img_channels', None),
                          label_dim=getattr(net, 'label_dim', None), use_fp16=bool(getattr(net, 'use_fp16', False)),
                          sigma_data=float(getattr(net, 'sigma_data', 0.5)), bytes=os.path.getsize(path))
        print(name, meta[name]['bytes'], 'bytes', meta[name]['n_tensors'], 'tensors')
    with open(os.path.join(OUT, 'edm_snapshot.json'), 'w') as f:
        json.dump(meta, f, indent=1)


if __name__ == '__main__':
    main()

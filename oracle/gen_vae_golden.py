"""Generate tests/golden/ref_vae.npz with the REAL reference Decoder (models/ldm/modules/diffusionmodules/model.py), in this container:

    python oracle/gen_vae_golden.py

The reference `Decoder` (and a plain torch Conv2d for AutoencoderKL.post_quant_conv, whose class needs pytorch_lightning to import) is
loaded with the oracle's seeded parameters and run on seeded latents on CPU fp32; its outputs pin oracle/vae_oracle.py
(tests/test_oracle_golden.py::test_vae_oracle_matches_reference).  /root/reference is only imported, never copied.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference/diff-solvers-main')


def main():
    from models.ldm.modules.diffusionmodules.model import Decoder
    from oracle import vae_oracle as VO
    out = {}
    for name, res in (('tiny_vae', 8), ('wide_vae', 8)):
        P, cfg = VO.make_params(name, seed=0)
        nlev = len(cfg['ch_mult'])
        dec = Decoder(ch=cfg['ch'], out_ch=cfg['out_ch'], ch_mult=tuple(cfg['ch_mult']), num_res_blocks=cfg['num_res_blocks'], attn_resolutions=[],
                      dropout=0.0, resamp_with_conv=True, in_channels=3, resolution=res * 2 ** (nlev - 1), z_channels=cfg['z_channels']).eval()
        sd = {k[len('decoder.'):]: v for k, v in P.items() if k.startswith('decoder.')}
        missing, unexpected = dec.load_state_dict(sd, strict=True)
        pq = torch.nn.Conv2d(cfg['embed_dim'], cfg['z_channels'], 1)
        pq.load_state_dict({'weight': P['post_quant_conv.weight'], 'bias': P['post_quant_conv.bias']})
        g = torch.Generator().manual_seed(3)
        z = torch.randn(2, cfg['z_channels'], res, res, generator=g) * cfg['scale_factor'] * 1.3
        with torch.no_grad():
            x = dec(pq(z / cfg['scale_factor']))
        out[f'vae/{name}/z'] = z.numpy()
        out[f'vae/{name}/x'] = x.numpy()
        print(name, tuple(z.shape), '->', tuple(x.shape), 'max|x|', float(x.abs().max()))
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'ref_vae.npz'), **out)


if __name__ == '__main__':
    main()

"""Generate tests/golden/ref_clip.npz: outputs of Hugging Face transformers' own CLIPTextModel (the third-party module the reference's
FrozenCLIPEmbedder wraps, models/ldm/modules/encoders/modules.py:137-159) on the seeded parameters of oracle/clip_oracle.make_params.

    python oracle/gen_clip_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import clip_oracle as CO  # noqa: E402


def main():
    import transformers
    from transformers import CLIPTextConfig, CLIPTextModel
    G = {}
    for name in ('tiny_clip',):
        P, cfg = CO.make_params(name, seed=0)
        hf = CLIPTextModel(CLIPTextConfig(hidden_act='quick_gelu', layer_norm_eps=1e-5, **cfg)).eval().requires_grad_(False)
        sd = hf.state_dict()
        assert list(sd.keys()) == list(P.keys()), [k for k in sd if k not in P][:4] + [k for k in P if k not in sd][:4]
        hf.load_state_dict(P)
        g = torch.Generator().manual_seed(11)
        ids = torch.randint(0, cfg['vocab_size'], (3, cfg['max_position_embeddings']), generator=g)
        ids[:, 0] = cfg['vocab_size'] - 2                    # <|startoftext|>-like
        ids[0, 20:] = cfg['vocab_size'] - 1                  # padded with <|endoftext|> as the tokenizer does (modules.py:153-154)
        with torch.no_grad():
            out = hf(input_ids=ids).last_hidden_state
        G[f'{name}/ids'] = ids.numpy()
        G[f'{name}/out'] = out.numpy()
    G['transformers_version'] = np.frombuffer(transformers.__version__.encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'ref_clip.npz'), **G)
    print('wrote ref_clip.npz', {k: v.shape for k, v in G.items()})


if __name__ == '__main__':
    main()

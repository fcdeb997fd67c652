"""ORACLE — test infrastructure, not product code (see oracle/edm_oracle.py header).

CPU restatement of the text encoder behind `get_learned_conditioning` (diff-solvers-main/sample.py:286-289):
`FrozenCLIPEmbedder.forward` (models/ldm/modules/encoders/modules.py:137-159) returns
`CLIPTextModel(input_ids=tokens).last_hidden_state`.  The arithmetic lives in a third-party dependency of the reference, Hugging Face
`transformers` (`from transformers import CLIPTokenizer, CLIPTextModel`, modules.py:6; the reference pins no version -- this image has
transformers 5.5, whose CLIP text model is the architecture of `openai/clip-vit-large-patch14` as published: token + position
embedding, 12 pre-LayerNorm transformer layers with causal self-attention (scale head_dim^-0.5) and a quick-GELU MLP, final LayerNorm).
Pinned by tests/golden/ref_clip.npz: outputs of transformers' own CLIPTextModel run in this container on this file's parameter recipe
(oracle/gen_clip_golden.py).  The tokenizer (vocabulary files, fetched from the hub) is caller-side: the encoder starts at token ids.
"""
from collections import OrderedDict

import torch
import torch.nn.functional as F

CONFIGS = {
    # openai/clip-vit-large-patch14 text tower (what Stable Diffusion v1.x conditions on)
    'clip_l': dict(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, max_position_embeddings=77),
    # reduced net with the same structure (64-wide heads, 77 positions)
    'tiny_clip': dict(vocab_size=1000, hidden_size=128, intermediate_size=512, num_hidden_layers=2, num_attention_heads=2, max_position_embeddings=77),
}


def param_shapes(cfg):
    """Ordered (name -> shape) of CLIPTextModel.state_dict() (transformers/models/clip/modeling_clip.py)."""
    H, I, V, T = cfg['hidden_size'], cfg['intermediate_size'], cfg['vocab_size'], cfg['max_position_embeddings']
    sh = OrderedDict()
    sh['text_model.embeddings.token_embedding.weight'] = (V, H)
    sh['text_model.embeddings.position_embedding.weight'] = (T, H)
    for i in range(cfg['num_hidden_layers']):
        p = f'text_model.encoder.layers.{i}.'
        for n in ('k_proj', 'v_proj', 'q_proj', 'out_proj'):
            sh[p + f'self_attn.{n}.weight'] = (H, H)
            sh[p + f'self_attn.{n}.bias'] = (H,)
        sh[p + 'layer_norm1.weight'] = (H,)
        sh[p + 'layer_norm1.bias'] = (H,)
        sh[p + 'mlp.fc1.weight'] = (I, H)
        sh[p + 'mlp.fc1.bias'] = (I,)
        sh[p + 'mlp.fc2.weight'] = (H, I)
        sh[p + 'mlp.fc2.bias'] = (H,)
        sh[p + 'layer_norm2.weight'] = (H,)
        sh[p + 'layer_norm2.bias'] = (H,)
    sh['text_model.final_layer_norm.weight'] = (H,)
    sh['text_model.final_layer_norm.bias'] = (H,)
    return sh


def make_params(name, seed=0):
    """Seeded parameters with O(1) activations through the stack (normal weights scaled by fan_in^-1/2, LayerNorm gains near 1)."""
    cfg = dict(CONFIGS[name])
    g = torch.Generator().manual_seed(seed + 4321)
    P = OrderedDict()
    for k, shp in param_shapes(cfg).items():
        if 'embedding' in k:
            P[k] = torch.randn(shp, generator=g) * 0.5
        elif len(shp) == 2:
            P[k] = torch.randn(shp, generator=g) * shp[1] ** -0.5
        elif 'layer_norm' in k and k.endswith('.weight'):
            P[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        else:
            P[k] = 0.1 * torch.randn(shp, generator=g)
    return P, cfg


def text_forward(P, cfg, input_ids, taps=None):
    """CLIPTextTransformer.forward -> last_hidden_state [B, T, H] (modeling_clip.py: CLIPTextEmbeddings, CLIPEncoderLayer, CLIPAttention with
    a causal mask, CLIPMLP with quick_gelu = x * sigmoid(1.702 x), final_layer_norm)."""
    B, T = input_ids.shape
    H, nh = cfg['hidden_size'], cfg['num_attention_heads']
    d = H // nh
    eps = cfg.get('layer_norm_eps', 1e-5)
    h = P['text_model.embeddings.token_embedding.weight'][input_ids] + P['text_model.embeddings.position_embedding.weight'][:T][None]
    mask = torch.full((T, T), float('-inf')).triu(1)
    for i in range(cfg['num_hidden_layers']):
        p = f'text_model.encoder.layers.{i}.'
        lin = lambda n, x: F.linear(x, P[p + n + '.weight'], P[p + n + '.bias'])
        y = F.layer_norm(h, (H,), P[p + 'layer_norm1.weight'], P[p + 'layer_norm1.bias'], eps)
        q, k, v = (lin('self_attn.' + n, y).reshape(B, T, nh, d).transpose(1, 2) for n in ('q_proj', 'k_proj', 'v_proj'))
        w = torch.softmax(q @ k.transpose(-1, -2) * d ** -0.5 + mask, dim=-1)
        a = (w @ v).transpose(1, 2).reshape(B, T, H)
        h = h + lin('self_attn.out_proj', a)
        y = F.layer_norm(h, (H,), P[p + 'layer_norm2.weight'], P[p + 'layer_norm2.bias'], eps)
        u = lin('mlp.fc1', y)
        h = h + lin('mlp.fc2', u * torch.sigmoid(1.702 * u))
        if taps is not None:
            taps[f'layer{i}'] = h
    return F.layer_norm(h, (H,), P['text_model.final_layer_norm.weight'], P['text_model.final_layer_norm.bias'], eps)

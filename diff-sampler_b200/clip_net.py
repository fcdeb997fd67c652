"""B200CLIPTextEncoder — native text encoder of `get_learned_conditioning` (diff-solvers-main/sample.py:286-289;
models/ldm/modules/encoders/modules.py:137-159 `FrozenCLIPEmbedder`): token ids -> CLIPTextModel.last_hidden_state [B, 77, 768], the
`context` / `uc` tensors the CFG denoiser consumes.  SURVEY section 8(f)3.

    enc = B200CLIPTextEncoder.from_reference(net.model.cond_stage_model)       # the reference's FrozenCLIPEmbedder
    c = enc.encode(["a photo of an astronaut riding a horse"] * B)            # = net.model.get_learned_conditioning(prompts)
    c = enc(tokens)                                                           # from token ids (int tensor [B, 77])

Tokenisation stays with the reference's own `CLIPTokenizer` object (a vocabulary lookup on the host); everything after it runs as one
plan on the GPU (clip_plan.py).  No CPU fallback.
"""
import ctypes as C
from collections import OrderedDict

import torch

from . import _cstructs as S
from . import _lib
from . import clip_plan


class B200CLIPTextEncoder:
    def __init__(self, params, num_heads=None, eps=1e-5, tokenizer=None, max_length=77, device='cuda'):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise _lib.DsError('B200CLIPTextEncoder needs a CUDA device (no CPU fallback)')
        self.lib = _lib.load()
        self.cfg = clip_plan.clip_config(params)
        self.num_heads = num_heads
        self.eps = float(eps)
        self.tokenizer = tokenizer
        self.max_length = int(max_length)
        self.wb = clip_plan.pack_clip_weights(params, self.cfg)
        blob = self.wb.bytes()
        self._wh = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.ds_weights_create(blob, len(blob), C.byref(self._wh)), 'ds_weights_create')
        self._plans = {}
        self.total_launches = 0

    @classmethod
    def from_reference(cls, module, **kw):
        """`module`: the reference's FrozenCLIPEmbedder (has `.transformer` = CLIPTextModel, `.tokenizer`, `.max_length`; modules.py:139-146),
        or a transformers CLIPTextModel itself."""
        tm = getattr(module, 'transformer', module)
        sd = OrderedDict((k, v) for k, v in tm.state_dict().items() if k.startswith('text_model.') and 'position_ids' not in k)
        conf = getattr(tm, 'config', None)
        if conf is not None:
            if getattr(conf, 'hidden_act', 'quick_gelu') != 'quick_gelu':
                raise ValueError(f'unsupported CLIP activation {conf.hidden_act!r} (the lowered MLP is quick_gelu)')
            kw.setdefault('num_heads', int(conf.num_attention_heads))
            kw.setdefault('eps', float(conf.layer_norm_eps))
        kw.setdefault('tokenizer', getattr(module, 'tokenizer', None))
        kw.setdefault('max_length', int(getattr(module, 'max_length', 77)))
        return cls(sd, **kw)

    def _plan(self, B, T):
        ent = self._plans.get((B, T))
        if ent is None:
            pl = clip_plan.compile_clip_plan(self.cfg, self.wb, B, T, num_heads=self.num_heads, eps=self.eps)
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.check(self.lib.ds_unet_create(self._wh, C.cast(pl.ops_array, C.c_void_p), pl.n_ops, C.sizeof(S.PlanOp), pl.arena_bytes,
                                                   C.byref(h)), 'ds_unet_create')
            ent = (h, pl)
            self._plans[(B, T)] = ent
        return ent

    def __call__(self, tokens, out=None):
        """tokens: integer tensor [B, T] on the CUDA device -> last_hidden_state [B, T, hidden] fp32.  A str / list of str is tokenised
        first, as FrozenCLIPEmbedder.forward(text) does."""
        if isinstance(tokens, (str, list, tuple)):
            return self.encode(tokens)
        if tokens.device.type != 'cuda':
            raise _lib.DsError('B200CLIPTextEncoder: token ids must live on the CUDA device (no CPU fallback)')
        if tokens.dim() != 2:
            raise ValueError(f'expected token ids [B, T], got {tuple(tokens.shape)}')
        ids = tokens.to(torch.int32).contiguous()
        B, T = ids.shape
        h, pl = self._plan(B, T)
        if out is None:
            out = torch.empty(B, T, self.cfg['hidden_size'], device=ids.device, dtype=torch.float32)
        io = (C.c_void_p * 6)(ids.data_ptr(), out.data_ptr(), None, None, None, None)
        stream = torch.cuda.current_stream(ids.device).cuda_stream
        _lib.check(self.lib.ds_unet_forward_io(h, io, 6, C.c_void_p(stream)), 'ds_unet_forward_io')
        self.total_launches += self.lib.ds_unet_last_launch_count(h)
        return out

    forward = __call__

    def encode(self, text):
        """FrozenCLIPEmbedder.encode (modules.py:148-159): tokenise with the reference's tokenizer (padding to max_length), then run."""
        if self.tokenizer is None:
            raise _lib.DsError('B200CLIPTextEncoder.encode needs the tokenizer of the reference module (from_reference keeps it)')
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True, return_overflowing_tokens=False,
                             padding='max_length', return_tensors='pt')
        return self(enc['input_ids'].to(self.device))

    def debug_read(self, B, T, name, numel, dtype=torch.float32):
        h, pl = self._plan(B, T)
        t = torch.empty(numel, dtype=dtype)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.ds_unet_debug_read(h, pl.arena_offsets[name], t.data_ptr(), t.numel() * t.element_size(), C.c_void_p(stream)),
                   'ds_unet_debug_read')
        return t

    def __del__(self):
        try:
            for h, _ in self._plans.values():
                self.lib.ds_unet_destroy(h)
            self.lib.ds_weights_destroy(self._wh)
        except Exception:
            pass

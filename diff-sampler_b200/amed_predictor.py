"""AMED predictor (inference): the 9k-parameter MLP that maps the U-Net bottleneck and (t_cur, t_next) to the per-sample
(r, scale_dir, scale_time).  Reference: amed-solver-main/training/networks.py:56-155.

On a CUDA device the whole predictor -- both time embeddings, the bottleneck MLP, the sigmoid heads and the geometric intermediate time
t_mid = t_next^r * t_cur^(1-r) (solvers_amed.py:119) -- is ONE kernel launch per sampling step (`predict_native` -> ds_amed_predict,
csrc/solver.cu) instead of ~25 ATen launches.  `forward` keeps the reference module's call contract in plain torch ops: it is what a
caller sees when it treats this object as the reference's AMED_predictor, and what the kernel is tested against."""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib


class AMEDPredictor(torch.nn.Module):
    def __init__(self, state_dict, scale_dir=0.0, scale_time=0.0, noise_channels=8):
        super().__init__()
        self.scale_dir, self.scale_time = float(scale_dir), float(scale_time)
        self.noise_channels = noise_channels
        for k, v in state_dict.items():
            self.register_buffer(k.replace('.', '__'), torch.as_tensor(v).clone().float())

    @classmethod
    def from_reference(cls, module):
        """Build from the reference's training.networks.AMED_predictor module (same state_dict names)."""
        inner = getattr(module, 'module', module)            # DistributedDataParallel wrapper (solvers_amed.py:30-33)
        return cls(inner.state_dict(), scale_dir=float(getattr(inner, 'scale_dir', 0.0) or 0.0),
                   scale_time=float(getattr(inner, 'scale_time', 0.0) or 0.0)).to(next(inner.parameters()).device)

    def pack(self, device):
        """One flat fp32 buffer in the order the kernel reads (csrc/solver.cu amed_predict_kernel) + its dims."""
        key = str(device)
        hit = getattr(self, '_packed', None)
        if hit is not None and hit[0] == key:
            return hit[1], hit[2]
        g = lambda n: getattr(self, n.replace('.', '__')).detach().float().reshape(-1)
        w0 = getattr(self, 'enc_layer0__weight')
        w1 = getattr(self, 'enc_layer1__weight')
        parts = [g('map_layer0.weight'), g('map_layer0.bias'), g('enc_layer0.weight'), g('enc_layer0.bias'), g('enc_layer1.weight'),
                 g('enc_layer1.bias'), g('fc_r.weight'), g('fc_r.bias')]
        has_dir = bool(self.scale_dir) and hasattr(self, 'fc_scale_dir__weight')
        has_time = bool(self.scale_time) and hasattr(self, 'fc_scale_time__weight')
        if has_dir:
            parts += [g('fc_scale_dir.weight'), g('fc_scale_dir.bias')]
        if has_time:
            parts += [g('fc_scale_time.weight'), g('fc_scale_time.bias')]
        buf = torch.cat(parts).to(device).contiguous()
        dims = (C.c_int * 6)(w0.shape[1], w0.shape[0], w1.shape[0], self.noise_channels, int(has_dir), int(has_time))
        assert getattr(self, 'fc_r__weight').shape[1] == w1.shape[0] + 2 * self.noise_channels
        self._packed = (key, buf, dims)
        return buf, dims

    def predict_native(self, bottleneck, t_cur, t_next, B):
        """(r, scale_dir, scale_time, t_mid) as rows of one [4, B] device tensor, one kernel launch.  bottleneck: [B, 8, 8] device
        tensor or None (analytical first step: zeros); t_cur / t_next: 0-d device tensors (views into t_steps)."""
        dev = t_cur.device
        if dev.type != 'cuda':
            raise _lib.DsError('AMEDPredictor.predict_native needs CUDA tensors (no CPU fallback)')
        lib = _lib.load()
        w, dims = self.pack(dev)
        out = torch.empty(4, B, device=dev, dtype=torch.float32)
        bp = None
        if bottleneck is not None:
            bt = bottleneck.reshape(B, -1).to(torch.float32).contiguous()
            assert bt.shape[1] == dims[0]
            bp = bt.data_ptr()
        tc = t_cur.reshape(1).to(torch.float32).contiguous()
        tn = t_next.reshape(1).to(torch.float32).contiguous()
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.ds_amed_predict(w.data_ptr(), dims, bp, tc.data_ptr(), tn.data_ptr(), self.scale_dir, self.scale_time, out.data_ptr(), B,
                                       stream), 'ds_amed_predict')
        from . import solver_utils
        solver_utils.LAUNCHES[0] += 1
        return out

    def _lin(self, name, x):
        y = x @ getattr(self, name + '__weight').t()
        b = getattr(self, name + '__bias', None)
        return y if b is None else y + b

    def _time_emb(self, t):
        half = self.noise_channels // 2
        freqs = torch.arange(half, dtype=torch.float32, device=t.device) / (half - 1)          # endpoint=True
        e = t.reshape(1,).float().ger((1 / 10000) ** freqs)
        e = torch.cat([e.sin(), e.cos()], dim=1)                                                 # [cos,sin] with the halves swapped
        return F.silu(self._lin('map_layer0', e))

    def forward(self, unet_bottleneck, t_cur, t_next, class_labels=None):
        B = unet_bottleneck.shape[0]
        emb = torch.cat([self._time_emb(t_cur).repeat(B, 1), self._time_emb(t_next).repeat(B, 1)], dim=1)
        z = self._lin('enc_layer1', F.silu(self._lin('enc_layer0', unet_bottleneck.reshape(B, -1))))
        out = torch.cat([z, emb], dim=1)
        r = torch.sigmoid(self._lin('fc_r', out))
        res = [r]
        if self.scale_dir:
            res.append(torch.sigmoid(self._lin('fc_scale_dir', out)) / (1 / (2 * self.scale_dir)) + (1 - self.scale_dir))
        if self.scale_time:
            res.append(torch.sigmoid(self._lin('fc_scale_time', out)) / (1 / (2 * self.scale_time)) + (1 - self.scale_time))
        return res[0] if len(res) == 1 else tuple(res)

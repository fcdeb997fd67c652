"""AMED predictor (inference): the 9k-parameter MLP that maps the U-Net bottleneck and (t_cur, t_next) to the per-sample
(r, scale_dir, scale_time).  Reference: amed-solver-main/training/networks.py:56-155.  It is a handful of tiny dense layers
on [B, 64] inputs — negligible next to the denoiser — and stays in PyTorch on the device."""
import torch
import torch.nn.functional as F


class AMEDPredictor(torch.nn.Module):
    def __init__(self, state_dict, scale_dir=0.0, scale_time=0.0, noise_channels=8):
        super().__init__()
        self.scale_dir, self.scale_time = float(scale_dir), float(scale_time)
        self.noise_channels = noise_channels
        for k, v in state_dict.items():
            self.register_buffer(k.replace('.', '__'), torch.as_tensor(v).clone().float())

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

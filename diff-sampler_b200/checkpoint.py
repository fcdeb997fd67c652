"""EDM checkpoint importer: `network-snapshot-*.pkl` -> parameter dict -> B200Net.

The reference loads its EDM networks with `pickle.load(f)['ema']` (sample.py:81-82).  Those pickles are written through
`torch_utils/persistence.py`: every `@persistent_class` module is reduced to `_reconstruct_persistent_obj(meta)` (:123-131) where
`meta` carries the *source code* of `networks_edm.py` plus the instance `__dict__`, and unpickling `exec`s that source (:222-234).
So the stock path needs `torch_utils` and `dnnlib` importable and runs code from the file.

This importer needs neither: a restricted `pickle.Unpickler` maps the persistence hook, `dnnlib.util.EasyDict` and the torch
container modules to inert record objects, accepts only tensor / ndarray / builtin reconstruction otherwise, never executes the
embedded source, and walks the `_parameters` / `_buffers` / `_modules` records into the same flat names `state_dict()` would give
(`model.enc.32x32_block0.conv0.weight`, ...).  `B200Net.from_pickle` feeds that to the plan compiler.

Only the preconditioner the hot path covers is accepted (`EDMPrecond`, networks_edm.py:459-500).  fp16 checkpoints
(ImageNet-64, `use_fp16=True`, :486) are imported as fp32 values: the kernels split every operand into fp16 planes themselves.
"""
import collections
import io
import pickle

import numpy as np
import torch

__all__ = ['load_edm_pickle', 'CheckpointError']


class CheckpointError(RuntimeError):
    pass


class _Record:
    """Stand-in for an unpickled object: keeps the class path and whatever state pickle hands over."""
    _path = '?'
    state = {}              # pickle creates instances with cls.__new__ (no __init__): objects without a BUILD state read this default

    def __init__(self, *args, **kwargs):
        self.args, self.kwargs, self.state = args, kwargs, {}

    def __setstate__(self, state):
        self.state = state if isinstance(state, dict) else {'__state__': state}

    def __repr__(self):
        return f'<record {self._path}>'


def _record_class(path):
    return type('Record_' + path.replace('.', '_'), (_Record,), {'_path': path})


class _EasyDict(dict):
    """dnnlib.util.EasyDict (dnnlib/util.py:38-50): a dict with attribute access; pickled as a plain dict subclass."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


class _Persistent(_Record):
    """What `_reconstruct_persistent_obj(meta)` stands for: class name + instance __dict__ (persistence.py:185-207)."""
    _path = 'persistent'


def _reconstruct_persistent_obj(meta):
    if meta.get('type') != 'class':
        raise CheckpointError(f"unsupported persistent record type {meta.get('type')!r}")
    r = _Persistent()
    r.class_name = meta['class_name']
    r.state = dict(meta['state']) if meta.get('state') is not None else {}
    return r                                     # meta['module_src'] is deliberately ignored: nothing from the file is executed


def _load_from_bytes(b):
    """torch.storage._load_from_bytes with the tensor-only loader (plain `pickle.dump` of tensors nests one torch.save per storage)."""
    return torch.load(io.BytesIO(b), map_location='cpu', weights_only=True)


try:                                            # numpy >= 2 moved the pickle helpers; old snapshots name numpy.core
    from numpy._core import multiarray as _np_ma
except ImportError:                             # pragma: no cover
    from numpy.core import multiarray as _np_ma

_ALLOWED = {
    ('torch_utils.persistence', '_reconstruct_persistent_obj'): _reconstruct_persistent_obj,
    ('dnnlib.util', 'EasyDict'): _EasyDict,
    ('collections', 'OrderedDict'): collections.OrderedDict,
    ('torch._utils', '_rebuild_tensor_v2'): torch._utils._rebuild_tensor_v2,
    ('torch._utils', '_rebuild_parameter'): torch._utils._rebuild_parameter,
    ('torch.storage', '_load_from_bytes'): _load_from_bytes,
    ('torch', 'Size'): torch.Size,
    ('torch', 'device'): torch.device,
    ('numpy', 'ndarray'): np.ndarray,
    ('numpy', 'dtype'): np.dtype,
}
for _m in ('numpy.core.multiarray', 'numpy._core.multiarray'):
    _ALLOWED[(_m, '_reconstruct')] = _np_ma._reconstruct
    _ALLOWED[(_m, 'scalar')] = _np_ma.scalar
for _n in ('FloatStorage', 'HalfStorage', 'DoubleStorage', 'LongStorage', 'IntStorage', 'BoolStorage', 'BFloat16Storage', 'ByteStorage'):
    _ALLOWED[('torch', _n)] = getattr(torch, _n)


class _Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        fn = _ALLOWED.get((module, name))
        if fn is not None:
            return fn
        if module.startswith('torch.nn.modules.') or module == 'torch.nn.parameter':
            if (module, name) == ('torch.nn.parameter', 'Parameter'):
                return torch.nn.Parameter
            return _record_class(module + '.' + name)          # ModuleDict / ModuleList containers: inert records of their __dict__
        raise CheckpointError(f'refusing to unpickle {module}.{name}: not part of an EDM network snapshot')


def _children(rec):
    mods = rec.state.get('_modules') or {}
    return mods.items()


def _flatten(rec, prefix, out):
    for k, v in (rec.state.get('_parameters') or {}).items():
        if v is not None:
            out[prefix + k] = v.detach()
    non_persistent = rec.state.get('_non_persistent_buffers_set') or set()
    for k, v in (rec.state.get('_buffers') or {}).items():
        if v is not None and k not in non_persistent:
            out[prefix + k] = v.detach()
    for k, child in _children(rec):
        if child is not None:
            _flatten(child, prefix + k + '.', out)


def load_edm_pickle(f, key='ema'):
    """Read an EDM network snapshot.  `f`: path or binary file object.  Returns (params, meta):
      params  OrderedDict name -> fp32 CPU tensor, named like `net.state_dict()` of the reference (resample filters dropped);
      meta    dict(img_resolution, img_channels, label_dim, sigma_min, sigma_max, sigma_data, use_fp16, class_name, model_type,
                   init_kwargs).
    Raises CheckpointError for anything that is not an EDMPrecond snapshot."""
    if isinstance(f, (str, bytes)) or hasattr(f, '__fspath__'):
        with open(f, 'rb') as fh:
            return load_edm_pickle(fh, key)
    try:
        top = _Unpickler(f).load()
    except pickle.UnpicklingError as e:
        raise CheckpointError(f'not a readable pickle: {e}')
    net = top.get(key) if isinstance(top, dict) else top
    if not isinstance(net, _Persistent):
        raise CheckpointError(f"no persistent network under key {key!r} (found {type(net).__name__})")
    if net.class_name != 'EDMPrecond':
        raise CheckpointError(f'{net.class_name}: only EDMPrecond snapshots are on the accelerated path (networks_edm.py:459-500)')
    flat = collections.OrderedDict()
    _flatten(net, '', flat)
    params = collections.OrderedDict((k, v.to(torch.float32).contiguous()) for k, v in flat.items() if 'resample_filter' not in k)
    st = net.state
    kw = dict(st.get('_init_kwargs') or {})
    model = (st.get('_modules') or {}).get('model')
    meta = dict(img_resolution=int(st['img_resolution']), img_channels=int(st['img_channels']), label_dim=int(st['label_dim']),
                sigma_min=float(st.get('sigma_min', 0.0)), sigma_max=float(st.get('sigma_max', float('inf'))),
                sigma_data=float(st.get('sigma_data', 0.5)), use_fp16=bool(st.get('use_fp16', False)), class_name=net.class_name,
                model_type=getattr(model, 'class_name', None), init_kwargs=kw)
    return params, meta

"""EDM snapshot importer (diff-sampler_b200/checkpoint.py) against fixtures written by the REAL reference classes through its own
torch_utils/persistence.py (oracle/gen_edm_pickle.py): same tensor names, order and values as the reference's state_dict()."""
import hashlib
import io
import json
import os
import pickle

import pytest
import torch

from diff_sampler_b200 import checkpoint as CK
from diff_sampler_b200 import edm_nets, plan as planner

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
META = json.load(open(os.path.join(GOLD, 'edm_snapshot.json')))


def digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().to(torch.float32).contiguous().numpy().tobytes())
    return h.hexdigest()


@pytest.mark.parametrize('name', ['song', 'adm_fp16', 'song64'])
def test_snapshot_matches_reference_state_dict(name):
    m = META[name]
    params, meta = CK.load_edm_pickle(os.path.join(GOLD, m['file']))
    assert len(params) == m['n_tensors'] and list(params)[:4] == m['keys_head']
    assert digest(params) == m['digest']                       # names, order and every value, bit for bit
    assert (meta['img_resolution'], meta['img_channels'], meta['label_dim']) == (m['img_resolution'], m['img_channels'], m['label_dim'])
    assert meta['use_fp16'] == m['use_fp16'] and meta['sigma_data'] == m['sigma_data'] and meta['class_name'] == 'EDMPrecond'
    assert all(v.dtype == torch.float32 and v.device.type == 'cpu' for v in params.values())
    # the imported dict drives the same spec inference / weight packing as a live reference module (B200Net.from_reference)
    spec = edm_nets.spec_from_params(params, meta['img_resolution'], meta['img_channels'], meta['label_dim'])
    assert spec.kind == ('song' if meta['model_type'] == 'SongUNet' else 'adm')
    assert len(spec.enc + spec.dec) > 0
    if name == 'song64':        # wide enough for the native kernels: the snapshot lowers to a plan (what B200Net.from_pickle runs on the GPU)
        wb, info = planner.pack_weights(spec, params)
        pl = planner.compile_plan(spec, wb, info, 4, 1, 0)
        assert pl.n_ops > 20 and pl.meta['n_gemm'] > 8


def test_file_object_and_missing_key():
    with open(os.path.join(GOLD, META['song']['file']), 'rb') as f:
        params, _ = CK.load_edm_pickle(f)
    assert digest(params) == META['song']['digest']
    with pytest.raises(CK.CheckpointError):
        CK.load_edm_pickle(os.path.join(GOLD, META['song']['file']), key='no_such_entry')


def test_other_network_classes_are_refused():
    with pytest.raises(CK.CheckpointError, match='EDMPrecond'):
        CK.load_edm_pickle(os.path.join(GOLD, META['bare_unet']['file']))


def test_nothing_outside_the_snapshot_vocabulary_is_unpickled():
    """The stock loader exec()s source embedded in the file (persistence.py:222-234); this one resolves a fixed set of names only."""
    class Evil:
        def __reduce__(self):
            return (os.system, ('echo pwned',))
    blob = pickle.dumps(dict(ema=Evil()))
    with pytest.raises(CK.CheckpointError, match='refusing'):
        CK.load_edm_pickle(io.BytesIO(blob))
    with pytest.raises(CK.CheckpointError):
        CK.load_edm_pickle(io.BytesIO(b'not a pickle at all'))


def test_torch1_era_tensor_records_are_readable():
    """The published EDM snapshots were written by torch 1.12: tensors appear as `_rebuild_tensor_v2(storage, ...)` whose storage is
    `torch.storage._load_from_bytes(<legacy non-zip torch.save bytes>)`.  The restricted unpickler resolves exactly that chain (with the
    tensor-only loader inside)."""
    import collections
    import warnings
    t = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        b = io.BytesIO()
        torch.save(t.storage(), b, _use_new_zipfile_serialization=False)
    raw = b.getvalue()

    class Storage:
        def __reduce__(self):
            return (torch.storage._load_from_bytes, (raw,))

    class Tensor:
        def __reduce__(self):
            return (torch._utils._rebuild_tensor_v2, (Storage(), 0, (3, 4), (4, 1), False, collections.OrderedDict()))

    obj = CK._Unpickler(io.BytesIO(pickle.dumps({'w': Tensor()}))).load()
    assert torch.equal(obj['w'], t)

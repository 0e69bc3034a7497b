"""N4: KeypointsDataset / DeviceLoader against the live reference's DataLoader order (golden fixture from
oracle/gen_golden.py `dataset`) and against a stock torch DataLoader over the same dataset object."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def _joints(tmp_path, **kw):
    from monoloco_b200 import synthetic
    path = str(tmp_path / 'joints.json')
    synthetic.make_joints_json(path, **kw)
    return path


def test_device_loader_reproduces_reference_batches(tmp_path):
    from monoloco_b200.train.datasets import KeypointsDataset, DeviceLoader
    g = json.load(open(os.path.join(GOLDEN, 'ref_dataset_order.json')))
    path = _joints(tmp_path, seed=g['joints_seed'])
    torch.manual_seed(g['seed'])
    loaders = {ph: DeviceLoader(KeypointsDataset(path, ph), g['bs'], shuffle=True, with_names=True) for ph in ('train', 'val')}
    for rec in g['order']:
        ds = loaders[rec['phase']].dataset
        got = []
        for inputs, labels, names, kps in loaders[rec['phase']]:
            ids = [int(v) for v in inputs[:, 0].tolist()]
            got.append(ids)
            assert [int(nm[:6]) for nm in names] == ids
            assert torch.equal(labels, ds.outputs_all[ids]) and torch.equal(kps, ds.kps_all[ids])
            assert inputs.dtype == labels.dtype == kps.dtype == torch.float32
        assert got == rec['batches'], (rec['epoch'], rec['phase'])
    ds = KeypointsDataset(path, 'val')
    assert {k: ds.get_cluster_annotations(k)[2] for k in g['clusters']} == g['clusters']
    assert abs(float(ds.inputs_all.double().sum()) - g['x_sum']) < 1e-6 * abs(g['x_sum']) and ds.get_version() == g['version']
    x, y, n = ds.get_cluster_annotations('20')
    assert x.shape == (n, 34) and y.shape == (n, 10) and x.dtype == y.dtype == torch.float32


def test_device_loader_equals_stock_dataloader_stereo(tmp_path):
    """Same dataset object under torch's DataLoader and under DeviceLoader, same seed: identical tensors batch by batch."""
    from torch.utils.data import DataLoader
    from monoloco_b200.train.datasets import KeypointsDataset, DeviceLoader
    ds = KeypointsDataset(_joints(tmp_path, n_train=301, n_val=40, stereo=True, seed=9), 'train')
    assert len(ds) == 301 and ds[3][0].shape == (68,) and ds[3][1].shape == (11,) and ds[3][2] == '000003.png'
    for bs, drop_last in ((64, False), (50, True), (512, False)):
        torch.manual_seed(3)
        ref = list(DataLoader(ds, batch_size=bs, shuffle=True, drop_last=drop_last))
        torch.manual_seed(3)
        dl = DeviceLoader(ds, bs, shuffle=True, drop_last=drop_last)
        mine = list(dl)
        assert len(mine) == len(ref) == len(dl)
        for (xi, yi, ni, ki), (xr, yr, nr, kr) in zip(mine, ref):
            assert torch.equal(xi, xr) and torch.equal(yi, yr) and torch.equal(ki, kr) and ni is None
        after = torch.rand(1)   # both consumed the global generator identically
        torch.manual_seed(3)
        list(DataLoader(ds, batch_size=bs, shuffle=True, drop_last=drop_last))
        assert torch.equal(after, torch.rand(1))
    seq = list(DeviceLoader(ds, 100, shuffle=False))
    assert torch.equal(torch.cat([b[0] for b in seq]), ds.inputs_all) and [b[0].shape[0] for b in seq] == [100, 100, 100, 1]

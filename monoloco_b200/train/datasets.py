"""
Dataset ingestion for the training step (reference: monoloco/train/datasets.py:43-97 KeypointsDataset and
its use through `DataLoader(KeypointsDataset(joints, phase), batch_size=bs, shuffle=True)`, trainer.py:102-107,150,209-232).

`KeypointsDataset` keeps the reference's attributes and methods, so it also works under a stock DataLoader.
`DeviceLoader` replaces the DataLoader for the fused step: the whole split (X [n,34|68], Y [n,10|11], kps [n,3,17]) moves to
the device ONCE; an epoch is one permutation and `n/bs` `index_select`s on the device -- no per-sample `__getitem__`,
no collation of Python tuples, no per-batch host->device copy.  With `shuffle=True` the permutation is drawn exactly the
way `DataLoader`/`RandomSampler` draw it (two int64 draws from the global torch generator per epoch: the iterator's base
seed, then the sampler seed that feeds `torch.randperm`), so under the same `torch.manual_seed` the batches are the
reference's batches, in the reference's order.
"""
import json

import numpy as np
import torch
from torch.utils.data import Dataset


def _f32(rows):
    return torch.from_numpy(np.asarray(rows, dtype=np.float32))


class KeypointsDataset(Dataset):
    """datasets.py:43-97: X / Y / names / kps of one phase of a joints JSON, plus the distance clusters used by evaluate()."""

    def __init__(self, joints, phase):
        assert phase in ['train', 'val', 'test']
        with open(joints, 'r') as f:
            dic_jo = json.load(f)
        self.inputs_all = _f32(dic_jo[phase]['X'])      # torch.tensor(list of floats) is float32 in the reference too
        self.outputs_all = _f32(dic_jo[phase]['Y'])
        self.names_all = dic_jo[phase]['names']
        self.kps_all = _f32(dic_jo[phase]['kps'])
        self.version = dic_jo['version']
        self.dic_clst = dic_jo[phase]['clst']

    def __len__(self):
        return self.inputs_all.shape[0]

    def __getitem__(self, idx):
        return self.inputs_all[idx, :], self.outputs_all[idx], self.names_all[idx], self.kps_all[idx, :]

    def get_cluster_annotations(self, clst):
        """Inputs / labels / count of one distance cluster ('10', '20', '30', '50', '>50', ...)."""
        ys = self.dic_clst[clst]['Y']
        return _f32(self.dic_clst[clst]['X']), _f32(ys), len(ys)

    def get_version(self):
        return self.version


class DeviceLoader:
    """Device-resident, DataLoader-order-compatible batch iterator over a KeypointsDataset.

    Yields `(inputs, labels, names, kps)` like the reference's loader (`for inputs, labels, _, _ in loader`, trainer.py:150);
    `names` is `None` unless `with_names=True` (the only per-batch Python work)."""

    def __init__(self, dataset, batch_size, shuffle=True, device=None, with_names=False, drop_last=False):
        self.dataset = dataset
        self.batch_size = int(batch_size)
        self.shuffle, self.with_names, self.drop_last = shuffle, with_names, drop_last
        self.device = torch.device(device) if device is not None else dataset.inputs_all.device
        self.inputs = dataset.inputs_all.to(self.device)
        self.labels = dataset.outputs_all.to(self.device)
        self.kps = dataset.kps_all.to(self.device)

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _permutation(self):
        n = len(self.dataset)
        if not self.shuffle:
            return torch.arange(n)
        # torch.utils.data: _BaseDataLoaderIter.__init__ draws the base seed, RandomSampler.__iter__ the sampler seed
        torch.empty((), dtype=torch.int64).random_()
        seed = int(torch.empty((), dtype=torch.int64).random_().item())
        g = torch.Generator()
        g.manual_seed(seed)
        return torch.randperm(n, generator=g)

    def __iter__(self):
        perm = self._permutation()
        perm_dev = perm.to(self.device, non_blocking=True)
        n, bs = perm.shape[0], self.batch_size
        stop = (n // bs) * bs if self.drop_last else n
        for i in range(0, stop, bs):
            idx = perm_dev[i:i + bs]
            names = [self.dataset.names_all[j] for j in perm[i:i + bs].tolist()] if self.with_names else None
            yield self.inputs.index_select(0, idx), self.labels.index_select(0, idx), names, self.kps.index_select(0, idx)

"""
Multi-GPU sharding of the inference path: one process per GPU, detections shard over ranks (weights are
replicated, 34 MB), and the per-detection outputs are all-gathered over NVLink.

Two gather modes:
  * 'nccl'  -- one `all_gather_into_tensor` of the [B_local, 20] output rows per step (the baseline);
  * 'fused' -- the forward kernel's decode epilogue stores every output row directly into every rank's gather
               buffer (cudaIpc-mapped peer memory over NVLink/NVSwitch), so the transfer overlaps the compute tile by
               tile; the only thing left after the kernel is a barrier.
The reference has no multi-GPU path; row layout of the gathered tensor is
[raw(out) | pad | x, y, z, d, bi, yaw_pred, yaw_orig, aux] with GATHER_LD = 20 floats.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L_


def shard_range(n_rows, world, rank):
    """Contiguous block partition; the first n_rows % world ranks get one extra row."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_sizes(n_rows, world):
    return [shard_range(n_rows, world, r)[1] - shard_range(n_rows, world, r)[0] for r in range(world)]


def pack_rows(raw, dec):
    """[B,out] + [B,8] -> [B, GATHER_LD] gather rows (same layout the fused epilogue writes)."""
    rows = torch.zeros((raw.shape[0], L_.GATHER_LD), dtype=torch.float32, device=raw.device)
    rows[:, :raw.shape[1]] = raw
    rows[:, L_.GATHER_DEC:L_.GATHER_DEC + 8] = dec
    return rows


def unpack_rows(rows, out_size):
    return rows[:, :out_size], rows[:, L_.GATHER_DEC:L_.GATHER_DEC + 8]


def all_gather_rows(local_rows, n_total, group=None):
    """Variable-size all-gather of row blocks (uneven shards are padded to the largest shard)."""
    world = dist.get_world_size(group)
    sizes = shard_sizes(n_total, world)
    mx = max(sizes)
    if all(s == mx for s in sizes):
        out = torch.empty((world * mx, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
        dist.all_gather_into_tensor(out, local_rows.contiguous(), group=group)
        return out
    padded = torch.zeros((mx, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    padded[:local_rows.shape[0]] = local_rows
    out = torch.empty((world * mx, local_rows.shape[1]), dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], dim=0)


class PeerGatherBuffer:
    """This rank's gather buffer + IPC mappings of every peer's buffer (mlb_ipc_*)."""

    def __init__(self, n_total_rows, device_index, group=None):
        self.lib = L_.lib()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.n_rows = n_total_rows
        self.device_index = device_index
        self.bytes = n_total_rows * L_.GATHER_LD * 4
        self.local = C.c_void_p()
        handle = C.create_string_buffer(L_.IPC_HANDLE_BYTES)
        L_.check(self.lib.mlb_ipc_alloc(device_index, self.bytes, C.byref(self.local), handle), 'mlb_ipc_alloc')
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs = []
        self._opened = []
        for r in range(self.world):
            if r == self.rank:
                self.ptrs.append(self.local.value)
            else:
                p = C.c_void_p()
                L_.check(self.lib.mlb_ipc_open(device_index, handles[r], C.byref(p)), 'mlb_ipc_open')
                self.ptrs.append(p.value)
                self._opened.append(p)
        dist.barrier(group=group)

    def tensor(self):
        """View of the local gather buffer as a [n_rows, GATHER_LD] CUDA tensor (zero-copy)."""
        iface = {'shape': (self.n_rows, L_.GATHER_LD), 'typestr': '<f4', 'data': (self.local.value, False), 'version': 3}

        class _Wrap:
            __cuda_array_interface__ = iface
        return torch.as_tensor(_Wrap(), device=torch.device('cuda', self.device_index))

    def close(self):
        for p in self._opened:
            self.lib.mlb_ipc_close(p)
        self._opened = []
        if self.local.value:
            self.lib.mlb_ipc_free(self.local)
            self.local = C.c_void_p()


class ShardedLoco:
    """Data-parallel forward over raw keypoints: every rank holds the full model and its shard of detections."""

    def __init__(self, engine, n_total_rows, mode='fused', group=None):
        assert mode in ('fused', 'nccl')
        self.eng, self.mode, self.group = engine, mode, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_total = n_total_rows
        self.start, self.stop = shard_range(n_total_rows, self.world, self.rank)
        self.buf = PeerGatherBuffer(n_total_rows, engine.index, group) if mode == 'fused' else None
        self._flag = torch.zeros(1, dtype=torch.float32, device=engine.device)

    def forward(self, kps_local, kk, rows_per_group=0):
        """kps_local: this rank's [stop-start, 3, 17] CUDA keypoints.  Returns the gathered [n_total, GATHER_LD] rows."""
        assert kps_local.shape[0] == self.stop - self.start
        if self.mode == 'fused':
            self.eng.forward(kps_local, kk=kk, kind=L_.IN_KPS, rows_per_group=rows_per_group,
                             gather_ptrs=self.buf.ptrs, gather_row0=self.start)
            # peer stores are complete when the kernel has retired; a 1-element all-reduce enqueued behind it on the
            # stream is the cross-rank barrier (no host synchronisation)
            dist.all_reduce(self._flag, group=self.group)
            return self.buf.tensor()
        out = self.eng.forward(kps_local, kk=kk, kind=L_.IN_KPS, rows_per_group=rows_per_group)
        return all_gather_rows(pack_rows(out['raw'], out['dec']), self.n_total, self.group)

    def close(self):
        if self.buf is not None:
            self.buf.close()

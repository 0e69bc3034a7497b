"""
Multi-GPU sharding of the inference path: one process per GPU, detections shard over ranks (weights are
replicated, 34 MB), and the per-detection outputs are all-gathered over NVLink.

Two gather modes:
  * 'nccl'  -- one `all_gather_into_tensor` of the [B_local, 20] output rows per step (the baseline);
  * 'fused' -- the forward kernel's decode epilogue stores every output row directly into every rank's gather
               buffer (cudaIpc-mapped peer memory over NVLink/NVSwitch), so the transfer overlaps the compute tile by
               tile, and the kernel's last CTA completes the exchange with a release/acquire flag protocol on the
               same peer memory: a step is one launch, no NCCL call in the data plane.
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
    """This rank's gather memory + IPC mappings of every peer's (mlb_ipc_*).  One allocation per rank:

        [ buffer 0: n_rows x GATHER_LD fp32 | buffer 1: same | completion flags: world x GATHER_FLAG_STRIDE uint32 ]

    The two row buffers alternate by step parity (a peer's step N+1 stores never land in the buffer this rank's
    consumers of step N are still reading); the flags carry the device-side completion protocol of mlb_forward
    (`gather_epoch`)."""

    def __init__(self, n_total_rows, device_index, group=None):
        self.lib = L_.lib()
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        if self.world > L_.MLB_MAX_PEERS:
            raise RuntimeError("fused all-gather supports up to %d ranks (one NVSwitch domain)" % L_.MLB_MAX_PEERS)
        self.n_rows = n_total_rows
        self.device_index = device_index
        self.buf_bytes = (max(n_total_rows, 1) * L_.GATHER_LD * 4 + 255) // 256 * 256
        self.flag_bytes = L_.MLB_MAX_PEERS * L_.GATHER_FLAG_STRIDE * 4
        self.bytes = 2 * self.buf_bytes + self.flag_bytes
        self.local = C.c_void_p()
        handle = C.create_string_buffer(L_.IPC_HANDLE_BYTES)
        L_.check(self.lib.mlb_ipc_alloc(device_index, self.bytes, C.byref(self.local), handle), 'mlb_ipc_alloc')
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.base = []
        self._opened = []
        for r in range(self.world):
            if r == self.rank:
                self.base.append(self.local.value)
            else:
                p = C.c_void_p()
                L_.check(self.lib.mlb_ipc_open(device_index, handles[r], C.byref(p)), 'mlb_ipc_open')
                self.base.append(p.value)
                self._opened.append(p)
        dist.barrier(group=group)  # every rank has mapped every buffer (and its zeroed flags) before the first step

    def row_ptrs(self, parity):
        return [b + parity * self.buf_bytes for b in self.base]

    def flag_ptrs(self):
        return [b + 2 * self.buf_bytes for b in self.base]

    def tensor(self, parity=0):
        """View of one local gather buffer as a [n_rows, GATHER_LD] CUDA tensor (zero-copy)."""
        iface = {'shape': (self.n_rows, L_.GATHER_LD), 'typestr': '<f4',
                 'data': (self.local.value + parity * self.buf_bytes, False), 'version': 3}

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
    """Data-parallel forward over raw keypoints: every rank holds the full model and its shard of detections.

    mode 'fused' (default): ONE kernel launch per step and no collective library in the data plane -- the decode
    epilogue stores each row into every rank's gather buffer over NVLink, and the launch's last CTA runs a flag
    protocol on the same peer-mapped memory (release-store of the step's epoch into every rank's flag array, acquire-spin
    on this rank's own), so the kernel retires exactly when the whole gathered tensor is complete on this GPU.
    mode 'nccl': forward, then `all_gather_into_tensor` (the A/B baseline).

    Lifetime of the returned tensor (fused mode): a zero-copy view of one of two alternating buffers; it stays valid
    until the next-but-one `forward` call on this object, for consumers enqueued on the same CUDA stream."""

    def __init__(self, engine, n_total_rows, mode='fused', group=None):
        assert mode in ('fused', 'nccl')
        self.eng, self.mode, self.group = engine, mode, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_total = n_total_rows
        self.start, self.stop = shard_range(n_total_rows, self.world, self.rank)
        self.buf = PeerGatherBuffer(n_total_rows, engine.index, group) if mode == 'fused' else None
        self.epoch = 0
        self._stage = None

    def forward(self, kps_local, kk, rows_per_group=0):
        """kps_local: this rank's [stop-start, 3, 17] CUDA keypoints.  Returns the gathered [n_total, GATHER_LD] rows."""
        assert kps_local.shape[0] == self.stop - self.start
        if self.mode == 'fused':
            self.epoch += 1
            if self.epoch & 0xFFFFFFFF == 0:  # 0 means "protocol off" in the ABI
                self.epoch += 1
            parity = self.epoch & 1
            self.eng.forward(kps_local, kk=kk, kind=L_.IN_KPS, rows_per_group=rows_per_group,
                             gather_ptrs=self.buf.row_ptrs(parity), gather_row0=self.start,
                             gather_flags=self.buf.flag_ptrs(), gather_rank=self.rank, gather_epoch=self.epoch)
            return self.buf.tensor(parity)
        out = self.eng.forward(kps_local, kk=kk, kind=L_.IN_KPS, rows_per_group=rows_per_group)
        return all_gather_rows(pack_rows(out['raw'], out['dec']), self.n_total, self.group)

    def forward_host(self, kps_local_host, kk, out_rows_host=None, rows_per_group=0):
        """Host buffers in, host buffers out (the end-to-end call): H2D of this rank's keypoints (pinned memory),
        sharded forward + all-gather, D2H of the whole gathered [n_total, GATHER_LD] tensor, stream sync."""
        dev = self.eng.device
        n_local = self.stop - self.start
        if self._stage is None or self._stage.shape[0] != n_local:
            self._stage = torch.empty((n_local, 3, 17), dtype=torch.float32, device=dev)
        self._stage.copy_(kps_local_host, non_blocking=True)
        rows = self.forward(self._stage, kk, rows_per_group=rows_per_group)
        if out_rows_host is None:
            out_rows_host = torch.empty((self.n_total, L_.GATHER_LD), dtype=torch.float32).pin_memory()
        out_rows_host.copy_(rows, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        self.eng.check_error()
        return out_rows_host

    def close(self):
        if self.buf is not None:
            self.buf.close()
            self.buf = None

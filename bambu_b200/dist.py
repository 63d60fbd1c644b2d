"""Multi-GPU plumbing: one process per GPU (torch.distributed), static sharding,
one gather.

The EM has no cross-group term (src/em.cpp has none) and samples are independent
(R/bambu.R:187), so the data path needs no collective: samples -- or, for a
single sample, gene groups -- are partitioned by a work estimate, each rank runs
its shard through the C-ABI, and the per-rank estimate tables are gathered to
rank 0, which is what ``combineCountSes`` does with ``cbind`` in the reference
(R/bambu_utilityFunctions.R:215-241).  NCCL is used for that gather only.
"""
from __future__ import annotations

import numpy as np


def lpt_partition(weights, n_parts):
    """Longest-processing-time-first: heaviest item to the currently lightest
    part.  Returns ``part_of_item`` (int array).  Deterministic (stable ties)."""
    w = np.asarray(weights, dtype=np.float64)
    order = np.argsort(-w, kind="stable")
    load = np.zeros(n_parts)
    part = np.empty(len(w), dtype=np.int64)
    for i in order:
        p = int(np.argmin(load))
        part[i] = p
        load[p] += w[i]
    return part


def shard_samples(sample_nnz, world):
    """Samples -> ranks by LPT on their entry counts."""
    part = lpt_partition(sample_nnz, world)
    return [np.nonzero(part == r)[0] for r in range(world)]


def shard_groups(arena, world, expected_iters=None):
    """One sample's gene groups -> ranks by LPT on nnz x expected iterations
    (giant loci first).  Returns, per rank, the group ids in arena order."""
    _, _, e = arena.group_shapes()
    return shard_groups_by_weight(e, world, expected_iters)


def shard_groups_by_weight(entries, world, expected_iters=None):
    """The same from per-group entry counts alone (e.g. the a-priori sizes of an annotation, before any rank has
    built its part of the arena): LPT places the heaviest groups -- the giant loci -- first."""
    w = np.asarray(entries, dtype=np.float64) * (1.0 if expected_iters is None else np.asarray(expected_iters, dtype=np.float64))
    part = lpt_partition(w, world)
    return [np.nonzero(part == r)[0] for r in range(world)]


class SharedHostTable:
    """The combined estimate table of a one-box run, gathered on the HOST side: one POSIX shared-memory segment
    that every rank maps (and page-locks when CUDA is there).  A rank's C-ABI call downloads its ``est[3, rows_r]``
    straight into its slice over its own PCIe link; ``gather()`` is then only a barrier, after which rank 0 holds
    every rank's table (``tables()``) -- the reference's ``combineCountSes`` cbind (R/bambu_utilityFunctions.R:215-241)
    without a second pass over any link.  ``EstimateGatherer`` is the NCCL alternative (tables travel GPU -> GPU 0 ->
    host over rank 0's link)."""

    def __init__(self, rows_local: int, group=None, pin: bool = True):
        import mmap
        import os
        import torch
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        allr = [None] * self.world
        dist.all_gather_object(allr, int(rows_local), group=group)
        self.rows = [int(x) for x in allr]
        off = np.concatenate(([0], np.cumsum([3 * 8 * r for r in self.rows]))).astype(np.int64)
        self.total = int(max(off[-1], 8))
        # a file in /dev/shm (tmpfs) mapped by every rank; rank 0 creates it, the name travels by broadcast
        name = [None]
        if self.rank == 0:
            try:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize < self.total + (64 << 20):
                    raise OSError("not enough room in /dev/shm")
                path = "/dev/shm/bambu_b200_%d_%x" % (os.getpid(), int.from_bytes(os.urandom(4), "little"))
                fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, self.total)
                os.close(fd)
                name[0] = path
            except OSError:
                name[0] = None
        dist.broadcast_object_list(name, src=0, group=group)
        if name[0] is None:        # every rank raises: the caller falls back to the NCCL gather
            raise RuntimeError("SharedHostTable: no shared memory for %d bytes" % self.total)
        self.path = name[0]
        fd = os.open(self.path, os.O_RDWR)
        self._map = mmap.mmap(fd, self.total)      # stays mapped until the process ends (numpy views point into it)
        os.close(fd)
        buf = np.frombuffer(self._map, dtype=np.uint8, count=self.total)
        self._views = [buf[off[r]:off[r + 1]].view(np.float64).reshape(3, self.rows[r]) for r in range(self.world)]
        self.mine = self._views[self.rank]
        self.mine[...] = 0.0                      # touch the pages from the rank that writes them
        self.pinned = False
        if pin and torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(buf.ctypes.data, self.total, 0)
            self.pinned = int(rc) == 0
        self._buf = buf
        dist.barrier(group=group)
        if self.rank == 0:                        # every rank has it mapped: the name can go
            try:
                os.unlink(self.path)
            except OSError:
                pass

    def gather(self):
        """Every rank has finished writing its slice when this returns."""
        self.dist.barrier(group=self.group)
        return self.tables()

    def tables(self):
        return self._views if self.rank == 0 else None

    def close(self):
        """Unpins the table.  The mapping itself lives until the process ends (callers may still hold views)."""
        import torch
        try:
            if self.pinned:
                torch.cuda.cudart().cudaHostUnregister(self._buf.ctypes.data)
                self.pinned = False
        except Exception:
            pass
        self.dist.barrier(group=self.group)


def gather_estimates(est, device=None, dst=0, group=None):
    """Gather every rank's ``est[3, rows_r]`` (host or device tensor) to ``dst``.
    Returns, on ``dst``, a list of host float64 arrays (one per rank, own shape);
    ``None`` elsewhere.  Ragged shards are padded to the longest for the collective.
    Backend-agnostic: NCCL moves device tensors, gloo (CPU tests) host tensors."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = est if isinstance(est, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(est))
    on_dev = device is not None and torch.device(device).type == "cuda"
    if on_dev:
        t = t.to(device, non_blocking=True)
    rows = torch.tensor([t.shape[1]], dtype=torch.int64, device=t.device)
    all_rows = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(all_rows, rows, group=group)
    all_rows = [int(x) for x in all_rows]
    mx = max(all_rows)
    if t.shape[1] < mx:
        pad = torch.zeros((3, mx), dtype=t.dtype, device=t.device)
        pad[:, :t.shape[1]] = t
        t = pad
    t = t.contiguous()
    bufs = [torch.empty_like(t) for _ in range(world)] if rank == dst else None
    dist.gather(t, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return [b[:, :n].cpu().numpy() for b, n in zip(bufs, all_rows)]


class EstimateGatherer:
    """Reusable gather of per-rank ``est[3, rows_r]`` tables to ``dst``: device staging
    buffers and the pinned host result are allocated once (cohort steps repeat the same
    shapes).  ``__call__`` takes this rank's pinned host table and returns, on ``dst``, a
    list of pinned host tensors (views, overwritten by the next call).

    ``overlap=True``: the upload of the table, the collective and ``dst``'s download are queued on a side
    stream and ``__call__`` returns at once, so the next step's EM runs while the previous step's tables
    travel; ``wait()`` (also called at the start of the next ``__call__``) completes them.  The caller must
    not overwrite the host table it passed before the next ``wait()`` -- alternate two tables."""

    def __init__(self, rows_local: int, device, dst: int = 0, group=None, overlap: bool = False):
        import torch
        import torch.distributed as dist
        self.dist, self.torch, self.group, self.dst = dist, torch, group, dst
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.device = torch.device(device)
        rows = torch.tensor([rows_local], dtype=torch.int64, device=self.device)
        allr = [torch.zeros_like(rows) for _ in range(self.world)]
        dist.all_gather(allr, rows, group=group)
        self.rows = [int(x) for x in allr]
        self.max_rows = max(self.rows)
        self.send = torch.zeros((3, self.max_rows), dtype=torch.float64, device=self.device)
        self.recv = self.host = None
        if self.rank == dst:
            self.recv = [torch.empty_like(self.send) for _ in range(self.world)]
            pin = self.device.type == "cuda"
            self.host = [torch.empty((3, n), dtype=torch.float64, pin_memory=pin) for n in self.rows]
        self.overlap = bool(overlap) and self.device.type == "cuda"
        self.stream = torch.cuda.Stream(self.device) if self.overlap else None
        self.pending = False

    def _enqueue(self, t):
        self.send[:, :t.shape[1]].copy_(t, non_blocking=True)
        self.dist.gather(self.send, self.recv, dst=self.dst, group=self.group)
        if self.rank == self.dst:
            for h, d, n in zip(self.host, self.recv, self.rows):
                h.copy_(d[:, :n], non_blocking=True)

    def wait(self):
        if self.pending:
            self.stream.synchronize()
            self.pending = False
        return self.host

    def __call__(self, est_host):
        t = est_host if isinstance(est_host, self.torch.Tensor) else self.torch.from_numpy(est_host)
        if self.overlap:
            self.wait()
            with self.torch.cuda.stream(self.stream):
                self._enqueue(t)
            self.pending = True
            return None
        self._enqueue(t)
        if self.rank != self.dst:
            return None
        if self.device.type == "cuda":
            self.torch.cuda.current_stream(self.device).synchronize()
        return self.host

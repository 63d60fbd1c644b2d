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
    """The combined estimate table of a one-box run, gathered on the HOST side: one file in /dev/shm that every rank
    maps.  A rank's C-ABI call downloads its ``est[3, rows_r]`` straight into its own slice (page-locked, over its own
    PCIe link) and then *publishes* the step by bumping its counter in the table's header; rank 0 *collects* a step by
    waiting until every counter has reached it -- after which it holds every rank's table, which is the reference's
    ``combineCountSes`` cbind (R/bambu_utilityFunctions.R:215-241) without a second pass over any link and without an
    all-rank barrier on the critical path (only rank 0 waits; the others go on with their next step).  The table is
    double-buffered: step k uses buffer k % 2, and a rank starting step k first waits until rank 0 has published its
    own step k - 1 (by then rank 0's caller has finished reading step k - 2, the buffer's previous content).  ``EstimateGatherer`` is the NCCL alternative (tables travel GPU -> GPU 0 -> host over rank 0's link).

    Layout: header of world + 1 cache lines (one int64 counter each: rank r's published step), then 2 buffers of all
    ranks' slices."""

    LINE = 64

    def __init__(self, rows_local: int = 0, group=None, pin: bool = True, local_shape=None):
        """``rows_local``: this rank's slice is ``est[3, rows_local]``; or ``local_shape``: any float64 block (e.g. the
        ``[4, samples, n_tx]`` assays of this rank's share of a cohort)."""
        import mmap
        import os
        import torch
        import torch.distributed as dist
        self.dist, self.group = dist, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        shape = tuple(int(x) for x in local_shape) if local_shape is not None else (3, int(rows_local))
        allr = [None] * self.world
        dist.all_gather_object(allr, shape, group=group)
        self.shapes = [tuple(x) for x in allr]
        self.rows = [x[-1] for x in self.shapes]
        sizes = [8 * int(np.prod(x)) for x in self.shapes]
        self.header = self.LINE * (self.world + 1)
        off = self.header + np.concatenate(([0], np.cumsum(sizes))).astype(np.int64)
        self.buf_bytes = int(sum(sizes))
        self.total = int(self.header + 2 * max(self.buf_bytes, 8))
        # rank 0 creates the file, the name travels by broadcast
        name = [None]
        if self.rank == 0:
            try:
                st = os.statvfs("/dev/shm")
                if st.f_bavail * st.f_frsize < self.total + (64 << 20):
                    raise OSError("not enough room in /dev/shm")
                path = "/dev/shm/bambu_b200_%d_%x" % (os.getpid(), int.from_bytes(os.urandom(4), "little"))
                fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                os.ftruncate(fd, self.total)      # zero-filled: every counter starts at 0
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
        self._buf = buf
        self._ctr = buf[:self.header].view(np.int64)             # counter of rank r at index r * LINE / 8
        self._views = [[buf[off[r] + b * self.buf_bytes: off[r + 1] + b * self.buf_bytes].view(np.float64).reshape(self.shapes[r])
                        for r in range(self.world)] for b in (0, 1)]
        self.step = 0                                            # steps this rank has published
        self._pinned = []
        for b in (0, 1):                                         # first touch + page-lock of this rank's own slices only
            mine = self._views[b][self.rank]
            mine[...] = 0.0
            if pin and torch.cuda.is_available() and mine.nbytes:
                rc = torch.cuda.cudart().cudaHostRegister(mine.ctypes.data, mine.nbytes, 0)
                if int(rc) == 0:
                    self._pinned.append(mine.ctypes.data)
        self.pinned = len(self._pinned) > 0
        dist.barrier(group=group)
        if self.rank == 0:                        # every rank has it mapped: the name can go
            try:
                os.unlink(self.path)
            except OSError:
                pass

    def _counter(self, r):
        return int(self._ctr[r * self.LINE // 8])

    def _spin(self, cond):
        import time
        n = 0
        while not cond():
            n += 1
            if n > 200:
                time.sleep(20e-6)

    @property
    def mine(self):
        """this rank's slice for its NEXT step; blocks while rank 0 still owns that buffer (step - 2 not collected)"""
        k = self.step + 1
        if k > 2 and self.rank != 0:
            # buffer k % 2 last held step k - 2; rank 0's caller is done reading it once rank 0 has published step k - 1
            self._spin(lambda: self._counter(0) >= k - 1)
        return self._views[k % 2][self.rank]

    def publish(self):
        """this rank's table of the step is complete in host memory"""
        self.step += 1
        self._ctr[self.rank * self.LINE // 8] = self.step

    def collect(self):
        """rank 0: wait for every rank's table of the step this rank has just published, mark it collected and return
        the per-rank views; other ranks return None at once"""
        if self.rank != 0:
            return None
        k = self.step
        self._spin(lambda: all(self._counter(r) >= k for r in range(self.world)))
        return self._views[k % 2]

    def gather(self):
        """publish + collect: the host-side gather of one step"""
        self.publish()
        return self.collect()

    def close(self):
        """Unpins the slices.  The mapping itself lives until the process ends (callers may still hold views)."""
        import torch
        for p in self._pinned:
            try:
                torch.cuda.cudart().cudaHostUnregister(p)
            except Exception:
                pass
        self._pinned = []
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


def cohort_assays(tables, sample_ids, n_samples_total, annotation_txid, incompatible_totals=None, emParameters=None,
                  sig_digit=5, quantify=None, group=None):
    """The quantification stage of a cohort on all ranks of one box: every rank quantifies ITS samples
    (``tables[k]`` is sample ``sample_ids[k]`` of the cohort; ``shard_samples`` makes such a split) with one
    ``bambu_quantDT_cohort`` call that writes the four assays of its samples straight into its page-locked slice of a
    shared host block; rank 0 then assembles the ``n_tx x n_samples_total`` matrices in cohort order -- ``bambu()``'s
    ``bplapply(readClassList, bambu.quantify)`` + ``combineCountSes`` (R/bambu.R:187-192).  Returns the dict of four
    matrices on rank 0, ``None`` elsewhere.  ``quantify(tables, annotation_txid, inc, emParameters, sig_digit, out)``
    defaults to ``quantify.bambu_quantDT_cohort_native`` (tests pass a stand-in)."""
    import torch.distributed as dist
    ann = np.ascontiguousarray(annotation_txid, dtype=np.int64)
    S, T = len(tables), len(ann)
    if len(sample_ids) != S:
        raise ValueError("one sample id per table")
    block = SharedHostTable(group=group, local_shape=(4, max(S, 1), max(T, 1)))
    out = block.mine
    inc = np.zeros(S) if incompatible_totals is None else np.asarray(incompatible_totals, dtype=np.float64)
    if quantify is None:
        from .quantify import bambu_quantDT_cohort_native

        def quantify(tables, ann, inc, par, sig, out):
            bambu_quantDT_cohort_native(tables, ann, inc, par, sig, out=out)
    if S:
        quantify(tables, ann, inc, emParameters, sig_digit, out)
    ids = [None] * dist.get_world_size(group)
    dist.all_gather_object(ids, [int(x) for x in sample_ids], group=group)
    views = block.gather()
    res = None
    if views is not None:
        names = ("counts", "CPM", "fullLengthCounts", "uniqueCounts")
        res = {n: np.full((T, int(n_samples_total)), np.nan, order="F") for n in names}
        for v, sid in zip(views, ids):
            for k, n in enumerate(names):
                if len(sid):
                    res[n][:, sid] = v[k, :len(sid), :T].T
    dist.barrier(group=group)         # rank 0 has copied: the block may go
    block.close()
    return res

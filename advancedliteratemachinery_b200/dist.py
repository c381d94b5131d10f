"""Data-parallel plumbing.  The path shards by page: every rank encodes / decodes its own
pages with replicated weights, so there are exactly two collectives -- ONE broadcast of the weights at start-up and ONE
gather of fixed-stride int32 sequence buffers per batch -- and none on the data path.

On the GPU both run inside libalm_ocr.so over its own NCCL communicator (`alm_comm_init`, `alm_broadcast_weights`,
`alm_gather_sequences`): the converted bf16 planes land in place on every rank, and a C++ host can do the same
without Python.  torch.distributed is only the rendezvous that ships the 128-byte NCCL id and the tensor shapes.  The
packing, padding and page re-ordering logic is transport-independent and is exercised on CPU with gloo
(tests/test_dist_cpu.py)."""
from __future__ import annotations

import threading
from collections import OrderedDict
from typing import Callable, List, Optional

import numpy as np
import torch
import torch.distributed as dist


def shard_pages(n_pages: int, rank: int, world: int) -> List[int]:
    """page i -> rank i mod world (round-robin)."""
    return list(range(rank, n_pages, world))


def pages_per_rank(n_pages: int, world: int) -> int:
    return (n_pages + world - 1) // world


# ------------------------------------------------------------------------------------------------ ordering
class CollectiveOrder:
    """Issue the collectives of several execution contexts (one NCCL communicator each, one host thread each) in ONE
    global order that is the same on every rank.

    NCCL's rule for several communicators in a process: their operations must be issued in the same order on all ranks.
    A gather kernel spins on the device until its peers arrive; if rank 0 issued context A's gather first and rank 1
    context B's, each rank would hold a spinning kernel (and possibly a cudaMalloc / hardware-queue slot behind it) that
    only the OTHER context of the peer can release -- a deadlock.  Work is numbered by a `ticket` that every rank derives
    the same way (the global step index); `run(ticket, fn)` blocks until all lower tickets have finished, then runs `fn`
    (the synchronous gather call).  A failure in any ticket releases the waiters with an error instead of hanging them."""

    def __init__(self, timeout_s: float = 300.0):
        self._cv = threading.Condition()
        self._next = 0
        self._failed: Optional[BaseException] = None
        self._timeout = timeout_s

    def reset(self, first_ticket: int = 0) -> None:
        with self._cv:
            self._next, self._failed = first_ticket, None

    def fail(self, exc: BaseException) -> None:
        with self._cv:
            if self._failed is None:
                self._failed = exc
            self._cv.notify_all()

    def run(self, ticket: int, fn: Callable):
        with self._cv:
            ok = self._cv.wait_for(lambda: self._failed is not None or self._next == ticket, timeout=self._timeout)
            if self._failed is not None:
                raise RuntimeError(f'collective {ticket} abandoned: an earlier one failed') from self._failed
            if not ok:
                self._failed = TimeoutError(f'collective {ticket} waited {self._timeout:.0f} s for ticket {self._next}')
                self._cv.notify_all()
                raise self._failed
        try:
            out = fn()
        except BaseException as e:
            self.fail(e)
            raise
        with self._cv:
            self._next = ticket + 1
            self._cv.notify_all()
        return out


class CollectiveLane:
    """The other way to keep one global order: ONE host thread per rank issues every collective of the job, in ticket
    order, on ONE dedicated context / communicator.  The compute threads only `submit(ticket, fn)` and go on with their next
    batch, so a rank that is momentarily ahead does not park its GPU work behind a peer's gather (with the synchronous form
    every batch is a rendezvous of all ranks, and the job runs at the pace of the slowest rank PER BATCH instead of on
    average).  `max_ahead` bounds how far the compute threads may run ahead of the lane."""

    def __init__(self, max_ahead: int = 16, timeout_s: float = 300.0, keep: int = 64):
        self._cv = threading.Condition()
        self._items, self._results = {}, {}
        self._next, self._stop = 0, True
        self._failed: Optional[BaseException] = None
        self._max_ahead, self._timeout, self._keep = max_ahead, timeout_s, keep
        self._thread: Optional[threading.Thread] = None

    def start(self, first_ticket: int = 0) -> None:
        self.stop()
        with self._cv:
            self._items.clear()
            self._results.clear()
            self._next, self._stop, self._failed = first_ticket, False, None
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self) -> None:
        with self._cv:
            self._stop = True
            self._cv.notify_all()
        if self._thread is not None:
            self._thread.join()
            self._thread = None

    def fail(self, exc: BaseException) -> None:
        with self._cv:
            if self._failed is None:
                self._failed = exc
            self._cv.notify_all()

    def submit(self, ticket: int, fn: Callable) -> None:
        with self._cv:
            ok = self._cv.wait_for(lambda: self._failed is not None or ticket - self._next < self._max_ahead,
                                   timeout=self._timeout)
            if self._failed is not None:
                raise RuntimeError(f'collective {ticket} not submitted: the lane failed') from self._failed
            if not ok:
                raise TimeoutError(f'collective {ticket}: the lane is stuck at ticket {self._next}')
            self._items[ticket] = fn
            self._cv.notify_all()

    def _run(self) -> None:
        while True:
            with self._cv:
                self._cv.wait_for(lambda: self._stop or self._failed is not None or self._next in self._items)
                if self._stop or self._failed is not None:
                    return
                ticket = self._next
                fn = self._items.pop(ticket)
            try:
                res = fn()
            except BaseException as e:   # noqa: BLE001 -- handed to the waiting threads
                self.fail(e)
                return
            with self._cv:
                self._results[ticket] = res
                self._results.pop(ticket - self._keep, None)
                self._next = ticket + 1
                self._cv.notify_all()

    def drain(self, upto: int) -> None:
        """Block until every ticket < `upto` has run."""
        with self._cv:
            ok = self._cv.wait_for(lambda: self._failed is not None or self._next >= upto, timeout=self._timeout)
            if self._failed is not None:
                raise RuntimeError('a collective failed') from self._failed
            if not ok:
                raise TimeoutError(f'collectives stuck at ticket {self._next} (waiting for {upto})')

    def result(self, ticket: int):
        with self._cv:
            return self._results.get(ticket)


# ------------------------------------------------------------------------------------------------ weights
def init_comm(ctx, device: Optional[torch.device] = None):
    """Create the library's NCCL communicator for `ctx`: rank 0 makes the id, torch.distributed ships its 128 bytes."""
    rank, world = dist.get_rank(), dist.get_world_size()
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, device=device)
    ctx.comm_init(box[0], rank, world)


def load_weights_broadcast(ctx, kind: int, state_dict: Optional[dict], src: int = 0,
                           device: Optional[torch.device] = None):
    """The weight start-up of a multi-GPU job: `src` converts the checkpoint (alm_load_weights), every other rank lays
    out placeholders of the same shapes, then ONE NCCL broadcast of the device slabs (bf16 hi/lo planes + fp32 vectors)
    puts the weights in place.  Needs `init_comm(ctx)` first.  Only names / shapes travel through torch.distributed."""
    rank = dist.get_rank()
    meta = [ctx.state_dict_meta(state_dict) if rank == src else None]
    dist.broadcast_object_list(meta, src=src, device=device)
    if rank == src:
        ctx.load_state_dict(kind, state_dict)
    else:
        ctx.load_placeholders(kind, meta[0])
    ctx.broadcast_weights(src)


def broadcast_state_dict(sd: Optional[dict], src: int = 0, device: Optional[torch.device] = None) -> OrderedDict:
    """Host-level alternative (any torch.distributed backend): one metadata broadcast + ONE flat fp32 broadcast of every
    tensor, returning the state dict on every rank.  The GPU job uses `load_weights_broadcast` instead."""
    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        meta[0] = [(k, tuple(v.shape), str(v.dtype)) for k, v in sd.items()]
    dist.broadcast_object_list(meta, src=src, device=device)
    meta = meta[0]
    total = sum(int(torch.Size(s).numel()) for _, s, _ in meta)
    dev = device or torch.device('cpu')
    if rank == src:
        flat = torch.cat([v.reshape(-1).to(torch.float32) for v in sd.values()]).to(dev)
    else:
        flat = torch.empty(total, dtype=torch.float32, device=dev)
    dist.broadcast(flat, src=src)
    flat = flat.cpu()
    out, off = OrderedDict(), 0
    for k, shape, dt in meta:
        n = int(torch.Size(shape).numel())
        t = flat[off:off + n].reshape(shape)
        out[k] = t.to(torch.int64) if dt == 'torch.int64' else t
        off += n
    return out


# ------------------------------------------------------------------------------------------------ sequences
def pack_sequences(outs, vocab, max_inst: int, rows: Optional[int] = None) -> torch.Tensor:
    """Per-rank decode results -> int32 [rows, 2, 1 + max_inst * (2 + 32 + L)] (count first, then pt|poly|rec rows;
    probabilities travel as their fp32 bit patterns in the second plane).  `rows` > len(outs) pads with empty pages so
    that every rank ships the same size when the pages do not divide evenly."""
    L = vocab.rec_length
    stride = 2 + 32 + L
    buf = torch.zeros(max(rows or 0, len(outs)), 2, 1 + max_inst * stride, dtype=torch.int32)
    for b, o in enumerate(outs):
        if o is None:
            continue
        (pt, poly, rec), (probs,) = o
        n = pt.numel() // 2
        buf[b, 0, 0] = n
        r = torch.cat([pt.reshape(n, 2), poly.reshape(n, 32), rec.reshape(n, L)], dim=1).to(torch.int32)
        buf[b, 0, 1:1 + n * stride] = r.reshape(-1)
        pr = torch.zeros(n, stride, dtype=torch.float32)
        pr[:, 34:] = probs
        buf[b, 1, 1:1 + n * stride] = pr.view(torch.int32).reshape(-1)
    return buf


def unpack_sequences(buf: torch.Tensor, vocab) -> list:
    L = vocab.rec_length
    stride = 2 + 32 + L
    outs = []
    for b in range(buf.shape[0]):
        n = int(buf[b, 0, 0])
        if n == 0:
            outs.append(None)
            continue
        rows = buf[b, 0, 1:1 + n * stride].reshape(n, stride).to(torch.int64)
        pr = buf[b, 1, 1:1 + n * stride].reshape(n, stride).contiguous().view(torch.float32)[:, 34:]
        outs.append(([rows[:, :2].reshape(1, -1), rows[:, 2:34].reshape(1, -1), rows[:, 34:][None]], [pr.clone()]))
    return outs


def interleave_pages(per_rank: List[list], n_pages: int) -> list:
    """per_rank[r][i] is global page r + i * world (round-robin sharding): back to page order, padding rows dropped."""
    world = len(per_rank)
    out = [None] * n_pages
    for r, rows in enumerate(per_rank):
        for i, o in enumerate(rows):
            p = r + i * world
            if p < n_pages:
                out[p] = o
    return out


def gather_sequences(outs, vocab, n_pages: Optional[int] = None, ctx=None, dst: Optional[int] = None,
                     device: Optional[torch.device] = None):
    """ONE gather of the packed sequences of a batch.

    `outs`: this rank's per-page results for its round-robin shard of `n_pages` global pages (default: every rank holds
    len(outs) pages).  Ranks may hold different page counts: buffers are padded to ceil(n_pages / world) rows.  Returns the
    results in GLOBAL PAGE ORDER (page p was decoded by rank p % world).
    Transport: `ctx` given -> the library's NCCL all-gather (alm_gather_sequences; every rank gets the list);
    otherwise torch.distributed (`dst` None -> all_gather, else gather to `dst`, None elsewhere)."""
    if ctx is not None:
        world, rank = ctx_world(ctx), None
    else:
        world, rank = dist.get_world_size(), dist.get_rank()
    if n_pages is None:
        n_pages = world * len(outs)
    rows = pages_per_rank(n_pages, world)
    assert len(outs) <= rows
    max_inst = max(1, vocab.pt_seq_length // 2)
    buf = pack_sequences(outs, vocab, max_inst, rows=rows)
    if ctx is not None:
        recv = torch.from_numpy(ctx.gather(buf.numpy(), world))
        return interleave_pages([unpack_sequences(recv[r], vocab) for r in range(world)], n_pages)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
    buf = buf.to(device)
    if dst is None:
        recv = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(recv, buf)
    else:
        recv = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
        dist.gather(buf, recv, dst=dst)
        if rank != dst:
            return None
    return interleave_pages([unpack_sequences(r.cpu(), vocab) for r in recv], n_pages)


def ctx_world(ctx) -> int:
    return int(getattr(ctx, 'comm_world', 1))

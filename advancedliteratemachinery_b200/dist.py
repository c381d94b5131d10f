"""Data-parallel plumbing over torch.distributed (NCCL on the GPU box, gloo in the CPU tests).

The path shards by page: every rank encodes/decodes its own pages with replicated weights, so there are
exactly two collectives: ONE broadcast of the packed weights at start-up and ONE gather
of fixed-stride int32 sequence buffers per batch.  No collective on the data path.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import List, Optional

import torch
import torch.distributed as dist


def shard_pages(n_pages: int, rank: int, world: int) -> List[int]:
    """page i -> rank i mod world (round-robin)."""
    return list(range(rank, n_pages, world))


def broadcast_state_dict(sd: Optional[dict], src: int = 0, device: Optional[torch.device] = None) -> OrderedDict:
    """One metadata broadcast (names/shapes, a few KB) + ONE flat fp32 broadcast of every tensor."""
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


def pack_sequences(outs, vocab, max_inst: int) -> torch.Tensor:
    """Per-rank decode results -> int32 [B, 1 + max_inst * (2 + 32 + L)] (count first, then pt|poly|rec rows;
    probabilities travel as their fp32 bit patterns in a second plane)."""
    L = vocab.rec_length
    stride = 2 + 32 + L
    buf = torch.zeros(len(outs), 2, 1 + max_inst * stride, dtype=torch.int32)
    for b, o in enumerate(outs):
        if o is None:
            continue
        (pt, poly, rec), (probs,) = o
        n = pt.numel() // 2
        buf[b, 0, 0] = n
        rows = torch.cat([pt.reshape(n, 2), poly.reshape(n, 32), rec.reshape(n, L)], dim=1).to(torch.int32)
        buf[b, 0, 1:1 + n * stride] = rows.reshape(-1)
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


def gather_sequences(outs, vocab, dst: int = 0, device: Optional[torch.device] = None):
    """ONE gather of the packed sequences to `dst`; returns the flat list of per-page results there
    (rank-major: pages of rank 0, then rank 1, ...), None elsewhere."""
    world, rank = dist.get_world_size(), dist.get_rank()
    max_inst = max(1, vocab.pt_seq_length // 2)
    buf = pack_sequences(outs, vocab, max_inst)
    if device is None:
        device = torch.device('cuda', torch.cuda.current_device()) if dist.get_backend() == 'nccl' else torch.device('cpu')
    buf = buf.to(device)
    recv = [torch.empty_like(buf) for _ in range(world)] if rank == dst else None
    dist.gather(buf, recv, dst=dst)
    if rank != dst:
        return None
    res = []
    for r in recv:
        res.extend(unpack_sequences(r.cpu(), vocab))
    return res

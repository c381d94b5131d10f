"""CPU (gloo, world_size 2): the two collectives of the data-parallel path -- weight broadcast and sequence
gather -- plus page sharding and the fixed-stride sequence packing."""
import os
import socket

import torch
import torch.multiprocessing as mp

from advancedliteratemachinery_b200.dist import (broadcast_state_dict, gather_sequences, interleave_pages, pack_sequences,
                                                 pages_per_rank, shard_pages, unpack_sequences)
from advancedliteratemachinery_b200.omniparser import OmniVocab


def _fake_out(n, L, seed):
    g = torch.Generator().manual_seed(seed)
    if n == 0:
        return None
    return ([torch.randint(0, 1000, (1, 2 * n), generator=g), torch.randint(0, 1000, (1, 32 * n), generator=g),
             torch.randint(1000, 1100, (1, n, L), generator=g)], [torch.rand(n, L, generator=g)])


def test_pack_unpack_roundtrip():
    v = OmniVocab(pt_seq_length=8)
    outs = [_fake_out(3, 25, 1), None, _fake_out(1, 25, 2), _fake_out(4, 25, 3)]
    back = unpack_sequences(pack_sequences(outs, v, 4), v)
    for a, b in zip(outs, back):
        assert (a is None) == (b is None)
        if a is not None:
            for x, y in zip(a[0], b[0]):
                assert torch.equal(x, y)
            assert torch.equal(a[1][0], b[1][0])


def test_round_robin_sharding_covers_every_page_once():
    for world in (1, 2, 4, 8):
        seen = sorted(p for r in range(world) for p in shard_pages(37, r, world))
        assert seen == list(range(37))


N_PAGES = 7  # does not divide by the world size: rank 0 holds 4 pages, rank 1 holds 3


def _page_out(p):
    return _fake_out(p % 3, 25, 100 + p)


def _same(a, b):
    if (a is None) != (b is None):
        return False
    return a is None or (all(torch.equal(x, y) for x, y in zip(a[0], b[0])) and torch.equal(a[1][0], b[1][0]))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    sd = None
    if rank == 0:
        g = torch.Generator().manual_seed(0)
        sd = {'a.weight': torch.randn(7, 5, generator=g), 'b.index': torch.arange(12).reshape(3, 4),
              'c.bias': torch.randn(9, generator=g)}
    got = broadcast_state_dict(sd, src=0)
    chk = float(got['a.weight'].sum() + got['c.bias'].sum()) + int(got['b.index'].sum())
    v = OmniVocab(pt_seq_length=8)
    outs = [_page_out(p) for p in shard_pages(N_PAGES, rank, world)]   # uneven shards
    exp = [_page_out(p) for p in range(N_PAGES)]
    res_all = gather_sequences(outs, v, n_pages=N_PAGES, device=torch.device('cpu'))            # all_gather: every rank
    ok = len(res_all) == N_PAGES and all(_same(a, b) for a, b in zip(exp, res_all))             # ... in PAGE order
    res_dst = gather_sequences(outs, v, n_pages=N_PAGES, dst=0, device=torch.device('cpu'))     # gather to rank 0
    if rank == 0:
        ok = ok and len(res_dst) == N_PAGES and all(_same(a, b) for a, b in zip(exp, res_dst))
    else:
        ok = ok and res_dst is None
    q.put((rank, chk, got['b.index'].dtype == torch.int64, ok))
    dist.destroy_process_group()


def test_uneven_shards_pad_and_come_back_in_page_order():
    v = OmniVocab(pt_seq_length=8)
    for world in (2, 3, 4, 8):
        rows = pages_per_rank(N_PAGES, world)
        per_rank = []
        for r in range(world):
            outs = [_page_out(p) for p in shard_pages(N_PAGES, r, world)]
            buf = pack_sequences(outs, v, 4, rows=rows)
            assert buf.shape[0] == rows                         # every rank ships the same size
            per_rank.append(unpack_sequences(buf, v))
        back = interleave_pages(per_rank, N_PAGES)
        assert len(back) == N_PAGES and all(_same(_page_out(p), back[p]) for p in range(N_PAGES))


def test_broadcast_and_gather_world2_gloo():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert abs(res[0][1] - res[1][1]) < 1e-6          # identical weights on both ranks
    assert all(r[2] and r[3] for r in res)

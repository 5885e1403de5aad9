"""Host logic of the row-sharded Sku table over a 2-rank `gloo` group on CPU: routing by owner,
compact-table order, id re-mapping and the gradient route back.  The owner-side row gather is a torch
indexing stub here (the CUDA gather kernel is exercised by the GPU tests); everything else is the product code.
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rows, dim, n, q):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        from cikm2020_dmt_b200.shard import RowExchange, RowShard, remap_ids
        g = torch.Generator().manual_seed(1234)
        full = torch.randn(rows, dim, generator=g)                      # identical on every rank
        shard = RowShard(rows, world, rank)
        local = full[shard.lo:shard.hi].clone()
        g2 = torch.Generator().manual_seed(77 + rank)
        need = torch.randint(-1, rows + 2, (n,), generator=g2)            # includes -1, rows, rows+1 = "no row"
        if rank == 0 and n >= 4:
            need[:4] = torch.tensor([0, rows - 1, shard.block - 1, shard.block])   # both sides of the boundary
        ex = RowExchange(shard, need)
        assert sum(ex.send_counts) == ex.n_valid
        compact = ex.fetch(lambda r: local[r.long()])
        valid = (need >= 0) & (need < rows)
        assert int(valid.sum()) == ex.n_valid
        assert torch.equal(ex.compact_row >= 0, valid)
        # every lookup finds its own row in the compact table
        assert torch.equal(compact[ex.compact_row[valid]], full[need[valid]])
        # compact rows are unique per lookup and sorted by variable row
        assert torch.equal(torch.sort(ex.compact_row[valid]).values, torch.arange(ex.n_valid))
        assert bool((ex.sorted_rows[1:] >= ex.sorted_rows[:-1]).all())
        # id re-mapping for the two lookup conventions
        zp = remap_ids(ex.compact_row, True, ex.n_valid)
        assert torch.equal(zp == 0, ~valid)
        raw = remap_ids(ex.compact_row, False, ex.n_valid)
        assert torch.equal(raw == ex.n_valid, ~valid)
        # gradient route: a gradient row that encodes its variable row must arrive at the owner of that row
        grads = (ex.sorted_rows.float()[:, None] + torch.arange(dim).float()[None, :] * 0.5)
        recv = ex.push_grads(grads)
        owned = ex.recv_rows + shard.lo
        assert recv.shape[0] == owned.numel() == sum(ex.recv_counts)
        assert bool(((owned >= shard.lo) & (owned < shard.hi)).all())
        assert torch.equal(recv, owned.float()[:, None] + torch.arange(dim).float()[None, :] * 0.5)
        # conservation: lookups sent == lookups received over the group
        tot = torch.tensor([ex.n_valid, recv.shape[0]])
        dist.all_reduce(tot)
        assert int(tot[0]) == int(tot[1])
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as exc:   # surface the failure in the parent
        import traceback
        q.put((rank, traceback.format_exc()))


@pytest.mark.parametrize("rows,n", [(1001, 5000), (7, 64), (64, 0)])
def test_row_exchange_world2_gloo(rows, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, rows, 8, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    for rank, msg in results:
        assert msg == "ok", "rank %d: %s" % (rank, msg)


def test_row_exchange_single_rank_is_identity_route():
    from cikm2020_dmt_b200.shard import RowExchange, RowShard
    full = torch.arange(40, dtype=torch.float32).view(10, 4)
    need = torch.tensor([3, 3, -1, 9, 0, 10])
    ex = RowExchange(RowShard(10, 1, 0), need)
    compact = ex.fetch(lambda r: full[r.long()])
    assert ex.n_valid == 4 and compact.shape == (4, 4)
    valid = torch.tensor([True, True, False, True, True, False])
    assert torch.equal(compact[ex.compact_row[valid]], full[need[valid]])
    assert ex.compact_row[2] == -1 and ex.compact_row[5] == -1


def test_row_shard_partition_covers_every_row_once():
    from cikm2020_dmt_b200.shard import RowShard
    for rows, world in [(5_000_000, 8), (10, 4), (3, 8), (1001, 2)]:
        seen = 0
        for r in range(world):
            sh = RowShard(rows, world, r)
            assert 0 <= sh.lo <= sh.hi <= rows
            seen += sh.local_rows
            if sh.local_rows:
                assert sh.owner(sh.lo) == r and sh.owner(sh.hi - 1) == r
        assert seen == rows

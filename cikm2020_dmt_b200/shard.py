"""Row-sharded embedding table over the data-parallel ranks ("Sku [V,32] row-sharded by contiguous
block, m/v sharded identically; fwd: all-to-all ids -> owner gather -> all-to-all rows; bwd: all-to-all
row-grads -> owner scatter-add + Adam on its shard").

The reference keeps the whole table on `/cpu:0` and every tower reads it (base.py:29,81-91); sharding is the
B200 replacement for that shared variable once it outgrows one GPU.

Per step and table, `RowExchange` routes every lookup of the local batch (a *variable row* index -- the
caller has already applied the zero-pad offset, so row i-1 and row i of the same id are simply two lookups and
need no halo) to the rank that owns the row and brings the rows back as a **compact table** whose row order is
the sorted lookup order.  The model kernels then run unchanged on the compact table with re-mapped ids; the
backward returns one gradient row per compact row, which travels back over the same route.

Only index arithmetic and `torch.distributed` collectives live here (NCCL on the GPU box, gloo in the CPU
tests); the row gather on the owner and the Adam update are the C-ABI kernels, passed in as callables.
"""
from typing import Callable, List, Optional

import torch
import torch.distributed as dist


class RowShard(object):
    """Contiguous block partition of `rows` variable rows over `world` ranks."""

    def __init__(self, rows: int, world: int, rank: int):
        self.rows, self.world, self.rank = int(rows), int(world), int(rank)
        self.block = (self.rows + self.world - 1) // self.world
        self.lo = min(self.rows, self.rank * self.block)
        self.hi = min(self.rows, self.lo + self.block)

    @property
    def local_rows(self):
        return self.hi - self.lo

    def owner(self, row):
        return row // self.block

    def bounds(self, device):
        b = torch.arange(self.world + 1, dtype=torch.int64, device=device) * self.block
        return torch.clamp(b, max=self.rows)


class RowExchange(object):
    """Routing of one step's lookups into one sharded table.

    need_rows   int tensor [n]: variable row of every lookup, < 0 or >= rows for "no row" (zero vector)
    After `fetch`, `compact_row[j]` is lookup j's row inside the compact table (or -1)."""

    def __init__(self, shard: RowShard, need_rows: torch.Tensor, group=None):
        self.shard, self.group = shard, group
        dev = need_rows.device
        rows = need_rows.to(torch.int64)
        rows = torch.where((rows < 0) | (rows >= shard.rows), torch.full_like(rows, -1), rows)
        sorted_rows, order = torch.sort(rows, stable=True)
        pos = torch.searchsorted(sorted_rows, shard.bounds(dev))          # [world+1]; invalid rows sort first
        world = shard.world
        if world > 1:
            all_pos = [torch.empty_like(pos) for _ in range(world)]
            dist.all_gather(all_pos, pos, group=group)
            all_pos = torch.stack(all_pos).cpu()                            # the one host sync of the exchange
        else:
            all_pos = pos[None].cpu()
        counts = (all_pos[:, 1:] - all_pos[:, :-1])                         # [src rank, dst rank]
        self.send_counts: List[int] = counts[shard.rank].tolist()
        self.recv_counts: List[int] = counts[:, shard.rank].tolist()
        self.n_invalid = int(all_pos[shard.rank, 0])
        self.n_valid = int(all_pos[shard.rank, -1]) - self.n_invalid
        self.sorted_rows = sorted_rows[self.n_invalid:self.n_invalid + self.n_valid].contiguous()
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel(), device=dev)
        self.compact_row = torch.where(rows >= 0, inv - self.n_invalid, torch.full_like(inv, -1))
        # owner side: the rows other ranks ask this rank for, as shard-local indices, grouped by source rank
        self.recv_rows = self._a2a(self.sorted_rows, self.send_counts, self.recv_counts) - shard.lo

    def _a2a(self, send: torch.Tensor, send_counts, recv_counts) -> torch.Tensor:
        out = torch.empty((sum(recv_counts),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        if self.shard.world == 1:
            out.copy_(send)
            return out
        dist.all_to_all_single(out, send.contiguous(), list(recv_counts), list(send_counts), group=self.group)
        return out

    def fetch(self, gather_local: Callable[[torch.Tensor], torch.Tensor]) -> torch.Tensor:
        """gather_local(shard-local int32 rows [m]) -> [m, D] rows of this rank's shard.  Returns the compact
        table [n_valid, D] (sorted-lookup order)."""
        served = gather_local(self.recv_rows.to(torch.int32))
        return self._a2a(served, self.recv_counts, self.send_counts)

    def push_grads(self, compact_grads: torch.Tensor) -> torch.Tensor:
        """compact_grads [n_valid, D] (one gradient row per compact row) -> the gradient rows of the lookups
        this rank owns, aligned with `recv_rows`."""
        return self._a2a(compact_grads, self.send_counts, self.recv_counts)


def remap_ids(compact_row: torch.Tensor, zero_pad: bool, n_valid: int) -> torch.Tensor:
    """Lookup ids into the compact table: with zero_pad the kernels read row id-1 and treat id 0 as the zero
    vector (base.py:87-89), so id = compact_row + 1 (0 when the lookup has no row); without it a lookup that
    has no row points one past the end, which the kernels read as zeros."""
    if zero_pad:
        return (compact_row + 1).to(torch.int32)
    return torch.where(compact_row < 0, torch.full_like(compact_row, n_valid), compact_row).to(torch.int32)

"""Multi-GPU plumbing of the count -> estimate path: one process per GPU, torch.distributed.

The path shards without a data-path collective up to the per-barcode histograms: every rank tallies (and, for UMI
data whose barcodes are rank-local, deduplicates) its own read range.  The exchange the north star names is that of
the per-barcode (k, n) histograms before the EM: a barcode's reads may sit in several read ranges, and the EM needs
the barcode's whole histogram.  The EM is partitioned by barcode (owner = barcode mod world): every rank sends its
sparse histogram rows to their owners (all-to-all; the all-gather form moves world times more) and the 8-byte
results are gathered back.  Works with NCCL (device tensors) and
with gloo (CPU tensors; used by the CPU tests for the exchange logic)."""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def all_gather_rows(rows: torch.Tensor, group=None) -> torch.Tensor:
    """All-gather of a [m, c] int64 tensor whose m differs per rank (padded to the largest m, then trimmed)."""
    world = dist.get_world_size(group)
    if world == 1:
        return rows
    m = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    sizes = [torch.zeros_like(m) for _ in range(world)]
    dist.all_gather(sizes, m, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(max(sizes), 1)
    padded = torch.zeros((cap, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    padded[:rows.shape[0]] = rows
    out = torch.empty((world * cap, rows.shape[1]), dtype=rows.dtype, device=rows.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    return torch.cat([out[r * cap:r * cap + sizes[r]] for r in range(world)], dim=0)


def owned_csr(rows: torch.Tensor, n_barcodes: int, rank: int, world: int):
    """rows = [barcode, k, n, count] (any order, duplicates allowed) -> the CSR triplets of the barcodes this rank
    owns (barcode % world == rank), barcodes ascending.  Returns (barcode ids, k int32, n int32, count int64, off int64)."""
    mine = rows[rows[:, 0] % world == rank] if world > 1 else rows
    order = torch.argsort(mine[:, 0], stable=True)
    mine = mine[order]
    owned = torch.arange(rank, n_barcodes, world, dtype=torch.int64, device=rows.device)
    local = torch.div(mine[:, 0] - rank, world, rounding_mode='floor').to(torch.int64)
    off = torch.zeros(owned.numel() + 1, dtype=torch.int64, device=rows.device)
    if mine.shape[0]:
        off[1:] = torch.cumsum(torch.bincount(local, minlength=owned.numel()), 0)
    return (owned, mine[:, 1].to(torch.int32).contiguous(), mine[:, 2].to(torch.int32).contiguous(),
            mine[:, 3].to(torch.int64).contiguous(), off)


def distributed_p_c(rows: torch.Tensor, p_e: torch.Tensor, n_barcodes: int,
                    em: Callable[[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                    group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """p_c of every barcode from rank-local histogram rows.
    rows: [m, 4] int64 or int32 (global barcode id, k, n, read count) of this rank's read range; p_e: [n_barcodes] float64
    (identical on every rank); em(k, n, count, off, p_e_owned) -> (p_c, status) for CSR groups -- the EM kernel
    (pipeline.em_pc) on a GPU.  Returns (p_c [n_barcodes] float64, status [n_barcodes] int32) on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    # every rank needs only the rows of the barcodes it owns: an all-to-all by owner moves 1/world of what the
    # all-gather of the histograms would (all_gather_rows stays available for callers that want every row everywhere)
    mine = exchange_by_owner(rows[:, 0] % world, rows, group=group)[0] if world > 1 else rows
    owned, k, n, c, off = owned_csr(mine, n_barcodes, rank, world)
    p_c_own, st_own = em(k, n, c, off, p_e[owned])
    if world == 1:
        return p_c_own, st_own
    per = (n_barcodes + world - 1) // world
    pad_pc = torch.full((per,), float('nan'), dtype=torch.float64, device=rows.device)
    pad_st = torch.full((per,), -1, dtype=torch.int32, device=rows.device)
    pad_pc[:owned.numel()] = p_c_own
    pad_st[:owned.numel()] = st_own
    g_pc = torch.empty(world * per, dtype=torch.float64, device=rows.device)
    g_st = torch.empty(world * per, dtype=torch.int32, device=rows.device)
    dist.all_gather_into_tensor(g_pc, pad_pc, group=group)
    dist.all_gather_into_tensor(g_st, pad_st, group=group)
    # rank r's slot j holds barcode r + j * world
    p_c = g_pc.view(world, per).t().reshape(-1)[:n_barcodes].contiguous()
    status = g_st.view(world, per).t().reshape(-1)[:n_barcodes].contiguous()
    return p_c, status


def read_range(n_reads: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous read range of a rank (read-range sharding of a coordinate-sorted BAM)."""
    per = (n_reads + world - 1) // world
    return min(rank * per, n_reads), min((rank + 1) * per, n_reads)


def exchange_by_owner(owner: torch.Tensor, *columns: torch.Tensor, group=None):
    """All-to-all of per-read rows to the rank that owns them (UMI groups may straddle read ranges, so
    before dedup every rank sends its fixed-size per-read records -- group key, 16 counters, priority fields -- to
    owner = barcode mod world).  `owner` [n] holds the destination rank of every local row; each column is a
    tensor with leading dimension n.  Returns the received columns, grouped by source rank in rank order (rows
    of one source keep their order, so a stable concatenation of read ranges is preserved)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return [c for c in columns]
    order = torch.argsort(owner, stable=True)
    send_counts = torch.bincount(owner, minlength=world)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    send_split = [int(x) for x in send_counts.tolist()]
    recv_split = [int(x) for x in recv_counts.tolist()]
    out = []
    for c in columns:
        src = c[order].contiguous()
        dst = torch.empty((sum(recv_split),) + tuple(c.shape[1:]), dtype=c.dtype, device=c.device)
        dist.all_to_all_single(dst, src, output_split_sizes=recv_split, input_split_sizes=send_split, group=group)
        out.append(dst)
    return out


def exchange_reads(ctx, barcode, umi, gene, score16, conv_n, strand, counts, world: int, transcriptome=None, group=None):
    """The per-read exchange before UMI dedup on device tensors, as ONE all-to-all: the local reads
    are partitioned by owner = barcode % world (dyn_owner_partition), their columns gathered into 40-byte rows
    (dyn_pack_reads), the rows exchanged (all_to_all_single over NCCL / NVLink), and unpacked on the receiving side
    (dyn_unpack_reads).  Rows of one source rank keep their order and sources arrive in rank order, so contiguous
    ascending read ranges stay in global read order -- what the "last row wins" tie rule of dedup needs.
    Returns the dict of received columns (see pipeline.unpack_reads)."""
    from . import pipeline
    perm, owner_counts = pipeline.owner_partition(ctx, barcode, world)
    rows = pipeline.pack_reads(ctx, perm, counts, barcode, umi, gene, score16, conv_n, strand, transcriptome)
    if world == 1:
        return pipeline.unpack_reads(ctx, rows)
    recv_counts = torch.empty_like(owner_counts)
    dist.all_to_all_single(recv_counts, owner_counts, group=group)
    send_split = [int(x) for x in owner_counts.tolist()]
    recv_split = [int(x) for x in recv_counts.tolist()]
    got = torch.empty((sum(recv_split), rows.shape[1]), dtype=rows.dtype, device=rows.device)
    dist.all_to_all_single(got, rows, output_split_sizes=recv_split, input_split_sizes=send_split, group=group)
    return pipeline.unpack_reads(ctx, got)


def exchange_kn_rows(ctx, keys, count32, n_barcodes: int, world: int, rank: int, group=None):
    """The histogram exchange on device tensors (no per-read exchange ran: every rank holds partial histograms of
    its read range).  keys: int64 (global barcode << 22 | k << 10 | n), count32: int32 reads per row.  Rows go to
    rank barcode % world in ONE partition + two all-to-alls, and the owner merges what it received
    (dyn_owner_partition, dyn_merge_kn, dyn_kn_csr).  Returns (owned barcode ids, k, n, count, off): the CSR
    histograms of the barcodes this rank owns, local barcode j = global barcode rank + j * world."""
    from . import pipeline
    if world > 1:
        perm, owner_counts = pipeline.owner_partition(ctx, (keys >> 22).to(torch.int32), world)
        order = perm.to(torch.int64)
        recv_counts = torch.empty_like(owner_counts)
        dist.all_to_all_single(recv_counts, owner_counts, group=group)
        send_split = [int(x) for x in owner_counts.tolist()]
        recv_split = [int(x) for x in recv_counts.tolist()]
        got_k = torch.empty(sum(recv_split), dtype=keys.dtype, device=keys.device)
        got_c = torch.empty(sum(recv_split), dtype=count32.dtype, device=keys.device)
        dist.all_to_all_single(got_k, keys[order].contiguous(), output_split_sizes=recv_split, input_split_sizes=send_split, group=group)
        dist.all_to_all_single(got_c, count32[order].contiguous(), output_split_sizes=recv_split, input_split_sizes=send_split,
                               group=group)
        keys, count32 = got_k, got_c
    owned = torch.arange(rank, n_barcodes, world, dtype=torch.int64, device=keys.device)
    mk, mc, n_rows = pipeline.merge_kn(ctx, keys, count32, world)
    k, n, c, off, _total = pipeline.kn_csr(ctx, mk, mc, n_rows, owned.numel())
    m = int(off[-1].item())
    return owned, k[:m], n[:m], c[:m], off


def distributed_p_c_keys(ctx, keys, count32, p_e, n_barcodes: int, em, group=None):
    """distributed_p_c for key-encoded rows, with the exchange and the CSR build on the device (exchange_kn_rows)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    owned, k, n, c, off = exchange_kn_rows(ctx, keys, count32, n_barcodes, world, rank, group=group)
    p_c_own, st_own = em(k, n, c, off, p_e[owned])
    if world == 1:
        return p_c_own, st_own
    per = (n_barcodes + world - 1) // world
    pad_pc = torch.full((per,), float('nan'), dtype=torch.float64, device=keys.device)
    pad_st = torch.full((per,), -1, dtype=torch.int32, device=keys.device)
    pad_pc[:owned.numel()] = p_c_own
    pad_st[:owned.numel()] = st_own
    g_pc = torch.empty(world * per, dtype=torch.float64, device=keys.device)
    g_st = torch.empty(world * per, dtype=torch.int32, device=keys.device)
    dist.all_gather_into_tensor(g_pc, pad_pc, group=group)
    dist.all_gather_into_tensor(g_st, pad_st, group=group)
    return (g_pc.view(world, per).t().reshape(-1)[:n_barcodes].contiguous(),
            g_st.view(world, per).t().reshape(-1)[:n_barcodes].contiguous())

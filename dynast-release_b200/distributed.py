"""Multi-GPU plumbing of the count -> estimate path: one process per GPU, torch.distributed.

The path shards without a data-path collective up to the per-barcode histograms: every rank tallies (and, for UMI
data whose barcodes are rank-local, deduplicates) its own read range.  The exchange the north star names is that of
the per-barcode (k, n) histograms before the EM: a barcode's reads may sit in several read ranges, and the EM needs
the barcode's whole histogram.  The EM is partitioned by barcode (owner = barcode mod world): every rank sends its
sparse histogram rows to their owners (all-to-all; the all-gather form moves world times more) and the 8-byte
results are gathered back.  Works with NCCL (device tensors) and
with gloo (CPU tensors; used by the CPU tests for the exchange logic)."""
from typing import Callable, Optional, Tuple

import numpy as np
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


PUSH_TILE = 256          # DYN_PUSH_TILE
PUSH_MAX_WORLD = 16      # DYN_PUSH_MAX_WORLD


class PeerExchange:
    """The per-read exchange with the transfer fused into the gather kernel: every rank's receive buffer is a cudaMalloc'ed
    block whose cudaIpc handle the other ranks of the node open, and dyn_push_reads writes each read's 40-byte row straight
    into the buffer of rank barcode % world over NVLink -- no permutation, no send buffer, no all-to-all on the data path.
    Protocol of one exchange (all on the caller's stream): dyn_push_plan -> all_gather of the per-owner counts (tiny; a rank
    passes it only after every rank has finished reading its receive buffer of the previous exchange) -> dyn_push_reads ->
    barrier (a one-element all_reduce: completes after every rank's push kernel has) -> dyn_unpack_reads.
    One instance per (context, process group); the buffers grow collectively when an exchange needs more room."""

    _instances = {}

    @classmethod
    def get(cls, ctx, group=None):
        key = (id(ctx), id(group))
        if key not in cls._instances:
            cls._instances[key] = cls(ctx, group)
        return cls._instances[key]

    def __init__(self, ctx, group=None):
        self.ctx, self.group = ctx, group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.capacity = 0
        self.mine = None
        self.peers = [None] * self.world
        self.usable = self.world <= PUSH_MAX_WORLD

    def _release(self):
        import ctypes as C
        for r, p in enumerate(self.peers):
            if p and r != self.rank:
                self.ctx.lib.dyn_peer_close(self.ctx.handle, C.c_void_p(p))
        if self.mine:
            self.ctx.lib.dyn_peer_free(self.ctx.handle, C.c_void_p(self.mine))
        self.mine, self.peers, self.capacity = None, [None] * self.world, 0

    def _grow(self, rows: int, device):
        """Collective: every rank (re)allocates `rows` rows and opens everybody else's handle."""
        import ctypes as C
        from ._lib import check
        torch.cuda.current_stream().synchronize()
        dist.barrier(group=self.group)                 # nobody is still writing into a buffer that is about to go away
        self._release()
        p = C.c_void_p()
        handle = (C.c_uint8 * 64)()
        rc = self.ctx.lib.dyn_peer_alloc(self.ctx.handle, int(rows) * 40, C.byref(p), handle)
        ok = torch.tensor([1 if rc == 0 else 0], device=device)
        mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
        every = torch.empty((self.world, 64), dtype=torch.uint8, device=device)
        dist.all_gather_into_tensor(every, mine, group=self.group)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if not bool(ok.item()):
            if rc == 0:
                self.ctx.lib.dyn_peer_free(self.ctx.handle, p)
            self.usable = False
            return False
        self.mine = p.value
        handles = every.cpu().numpy()
        opened = 1
        for r in range(self.world):
            if r == self.rank:
                self.peers[r] = self.mine
                continue
            q = C.c_void_p()
            buf = (C.c_uint8 * 64)(*handles[r].tolist())
            if self.ctx.lib.dyn_peer_open(self.ctx.handle, buf, C.byref(q)) != 0:
                opened = 0
                break
            self.peers[r] = q.value
        ok = torch.tensor([opened], device=device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if not bool(ok.item()):
            self._release()
            self.usable = False
            return False
        self.capacity = int(rows)
        return True

    def exchange(self, barcode, umi, gene, score16, conv_n, strand, counts, transcriptome=None):
        """Returns the dict of received columns (pipeline.unpack_reads), or None when peer memory is not available on this
        node (the caller then takes the all-to-all path; the decision is collective)."""
        import ctypes as C
        from . import pipeline
        from ._lib import check, ptr
        if not self.usable:
            return None
        ctx, world, dev = self.ctx, self.world, barcode.device
        n = barcode.numel()
        n_tiles = (n + PUSH_TILE - 1) // PUSH_TILE
        tile_base = torch.empty(max(n_tiles * world, 1), dtype=torch.int32, device=dev)
        owner_counts = torch.empty(world, dtype=torch.int64, device=dev)
        check(ctx.lib.dyn_push_plan(ctx.handle, ptr(barcode), n, world, ptr(tile_base), ptr(owner_counts), pipeline._stream()))
        matrix = torch.empty((world, world), dtype=torch.int64, device=dev)          # [source, owner]
        dist.all_gather_into_tensor(matrix, owner_counts, group=self.group)
        m = matrix.cpu().numpy()
        need = int(m.sum(axis=0).max())
        if need > self.capacity and not self._grow(int(need * 1.25) + 1024, dev):
            return None
        starts = np.concatenate([[0], np.cumsum(m[self.rank])[:-1]])
        dst_base = (C.c_int64 * world)(*[int(m[:self.rank, o].sum()) - int(starts[o]) for o in range(world)])
        peers = (C.c_void_p * world)(*self.peers)
        check(ctx.lib.dyn_push_reads(ctx.handle, ptr(tile_base), n, world, ptr(counts), ptr(barcode), ptr(umi), ptr(gene), ptr(score16),
                                     ptr(conv_n), ptr(strand), ptr(transcriptome), peers, dst_base, pipeline._stream()))
        done = torch.zeros(1, dtype=torch.int32, device=dev)
        dist.all_reduce(done, group=self.group)        # on the stream, after the push kernel: every rank's rows have landed
        n_recv = int(m[:, self.rank].sum())
        rows = pipeline.DevicePointer(self.mine, (n_recv, 40), dev)
        return pipeline.unpack_reads(ctx, rows)


def exchange_reads(ctx, barcode, umi, gene, score16, conv_n, strand, counts, world: int, transcriptome=None, group=None,
                   mode=None):
    """The per-read exchange before UMI dedup on device tensors.  Rows of one source rank keep their order and
    sources arrive in rank order, so contiguous ascending read ranges stay in global read order -- what the "last row
    wins" tie rule of dedup needs.  Two transports with identical results:
      'push'  (default on CUDA tensors, one node): PeerExchange -- rows are stored into the owner's memory over NVLink by
              the kernel that gathers them;
      'nccl'  (DYN_EXCHANGE=nccl, CPU / gloo tests, or when peer memory cannot be mapped): the local reads are partitioned
              by owner (dyn_owner_partition), gathered into 40-byte rows (dyn_pack_reads), exchanged in ONE
              all_to_all_single and unpacked (dyn_unpack_reads).
    Returns the dict of received columns (see pipeline.unpack_reads)."""
    import os
    from . import pipeline
    mode = mode or os.environ.get('DYN_EXCHANGE', 'push')
    if world > 1 and mode == 'push' and barcode.is_cuda:
        got = PeerExchange.get(ctx, group).exchange(barcode, umi, gene, score16, conv_n, strand, counts, transcriptome)
        if got is not None:
            return got
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

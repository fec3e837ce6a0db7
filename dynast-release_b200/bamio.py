"""Python face of the host BAM ingest (csrc/bampack.cpp, dyn_bam_pack): decode a coordinate-sorted BAM into the
packed records the kernels read plus the per-unit metadata the CSV outputs need."""
import ctypes as C
from typing import List, Optional

import numpy as np

from . import _lib

_FIELDS = {'blob': (0, np.uint8), 'rec_off': (1, np.uint32), 'ref_id': (2, np.int32), 'pos': (3, np.int32),
           'flag': (4, np.uint16), 'hi': (5, np.int32), 'score': (6, np.int32), 'gx_assigned': (7, np.uint8),
           'paired': (8, np.uint8), 'ref_len': (9, np.int32)}
_FIELDS.update({'barcode_key': (11, np.uint64), 'umi_key': (12, np.uint64), 'gene_idx': (13, np.int32), 'rec_index': (14, np.int64),
                'raw': (30, np.uint8), 'raw_off': (31, np.uint32), 'raw_mate': (32, np.uint8), 'raw_mate_off': (33, np.uint32)})
KEEP_RAW, KEEP_DUPLICATES = 0x1, 0x2
_POOLS = {'read_id': (20, 21), 'barcode': (22, 23), 'umi': (24, 25), 'gene': (26, 27), 'ref_name': (28, 29)}


class PackedBam:
    """Result of dyn_bam_pack. Numeric fields are numpy views into library-owned memory (valid until close());
    string fields are decoded on demand."""

    def __init__(self, path: str, barcode_tag: Optional[str] = None, umi_tag: Optional[str] = None,
                 gene_tag: Optional[str] = 'GX', require_gene: bool = False, n_threads: int = 8, flags: int = 0):
        self.lib = _lib.load()
        h = C.c_void_p()
        enc = lambda s: s.encode() if s else None
        _lib.check(self.lib.dyn_bam_pack_ex(path.encode(), enc(barcode_tag), enc(umi_tag), enc(gene_tag), int(require_gene),
                                            int(n_threads), int(flags), C.byref(h)))
        self._adopt(h)

    def _adopt(self, h):
        self.handle = h
        self.n_units = int(self.lib.dyn_bam_n_units(h))
        self.n_records_seen = int(self.lib.dyn_bam_n_records_seen(h))
        self._cache = {}

    @classmethod
    def from_handle(cls, lib, handle) -> 'PackedBam':
        """A chunk handed out by dyn_bam_stream_next."""
        self = cls.__new__(cls)
        self.lib = lib
        self._adopt(handle)
        return self

    def _raw(self, field: int, dtype):
        nb = C.c_int64(0)
        p = self.lib.dyn_bam_field(self.handle, field, C.byref(nb))
        n = nb.value // np.dtype(dtype).itemsize
        if not p or n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nb.value,)).view(dtype)[:n]

    def __getattr__(self, name):
        if name in _FIELDS:
            f, dt = _FIELDS[name]
            return self._raw(f, dt)
        if name in _POOLS:
            if name not in self._cache:
                self._cache[name] = self.strings(name)
            return self._cache[name]
        raise AttributeError(name)

    def strings(self, name: str) -> np.ndarray:
        """object array of the unit's strings (read_id, barcode, umi, gene) or of the reference names."""
        fd, fo = _POOLS[name]
        data = self._raw(fd, np.uint8).tobytes()
        off = self._raw(fo, np.uint32)
        n = len(off) - 1
        out = np.empty(n, dtype=object)
        if n == 0:
            return out
        lens = np.diff(off.astype(np.int64))
        if lens.min() == lens.max() and lens[0] > 0 and int(off[0]) == 0 and data.isascii():
            # one width (barcodes, UMIs, most gene ids): a fixed-width view decodes them all at once
            width = int(lens[0])
            fixed = np.frombuffer(data, dtype=f'S{width}', count=n)
            if not (np.frombuffer(data, dtype=np.uint8, count=n * width).reshape(n, width)[:, -1] == 0).any():      # 'S' strips trailing NULs
                out[:] = fixed.astype(f'U{width}')
                return out
        if data.isascii():          # byte offsets are character offsets: one decode, then slices
            text = data.decode()
            o = off.tolist()
            out[:] = [text[a:b] for a, b in zip(o[:-1], o[1:])]
            return out
        for i in range(n):
            out[i] = data[off[i]:off[i + 1]].decode()
        return out

    @property
    def header_text(self) -> str:
        return self._raw(10, np.uint8).tobytes().decode(errors='replace')

    def close(self):
        if self.handle:
            self.lib.dyn_bam_result_free(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def decode_key(key: int, suffix: bool = True) -> str:
    """The string a 2-bit barcode / UMI key stands for (the inverse of the packer's encoding)."""
    key = int(key)
    if key == 0:
        return 'NA'
    well = (key >> 58) if suffix else 0
    k = key & ((1 << 58) - 1) if suffix else key
    out = []
    while k > 1:
        out.append('ACGT'[k & 3])
        k >>= 2
    return ''.join(reversed(out)) + (f'-{well - 1}' if well else '')


class BamStream:
    """dyn_bam_stream_*: block range `part` of `n_parts` of a coordinate-sorted BAM, chunk by chunk (each chunk a
    PackedBam; close it when done -- its page-locked buffer is reused).  The next chunks are inflated and packed by
    the library's threads while the caller works on the current one.  gene_ids: the gene tag becomes an index into
    this list and barcode / UMI become 2-bit keys (fields barcode_key, umi_key, gene_idx) instead of strings."""

    def __init__(self, path: str, barcode_tag: Optional[str] = None, umi_tag: Optional[str] = None,
                 gene_tag: Optional[str] = 'GX', require_gene: bool = False, n_threads: int = 8, part: int = 0,
                 n_parts: int = 1, chunk_bytes: int = 0, lean: bool = False, keys: bool = False,
                 gene_ids: Optional[List[str]] = None):
        self.lib = _lib.load()
        enc = lambda s: s.encode() if s else None
        self._gene_ids = None
        o = _lib.BamStreamOpts(barcode_tag=enc(barcode_tag), umi_tag=enc(umi_tag), gene_tag=enc(gene_tag),
                               require_gene=int(require_gene), n_threads=int(n_threads), part=int(part), n_parts=int(n_parts),
                               chunk_bytes=int(chunk_bytes), lean=int(lean), keys=int(keys or gene_ids is not None))
        if gene_ids is not None:
            self._gene_ids = (C.c_char_p * len(gene_ids))(*[g.encode() for g in gene_ids])
            o.gene_ids = C.cast(self._gene_ids, C.POINTER(C.c_char_p))
            o.n_gene_ids = len(gene_ids)
        h = C.c_void_p()
        _lib.check(self.lib.dyn_bam_stream_open(path.encode(), C.byref(o), C.byref(h)))
        self.handle = h

    def _raw(self, field: int, dtype):
        nb = C.c_int64(0)
        p = self.lib.dyn_bam_stream_field(self.handle, field, C.byref(nb))
        if not p or nb.value == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nb.value,)).view(dtype).copy()

    @property
    def ref_len(self) -> np.ndarray:
        return self._raw(9, np.int32)

    @property
    def header_text(self) -> str:
        return self._raw(10, np.uint8).tobytes().decode(errors='replace')

    @property
    def ref_name(self) -> np.ndarray:
        data, off = self._raw(28, np.uint8).tobytes(), self._raw(29, np.uint32)
        return np.array([data[off[i]:off[i + 1]].decode() for i in range(len(off) - 1)], dtype=object)

    def __iter__(self):
        while True:
            h = C.c_void_p()
            _lib.check(self.lib.dyn_bam_stream_next(self.handle, C.byref(h)))
            if not h:
                return
            yield PackedBam.from_handle(self.lib, h)

    def close(self):
        if self.handle:
            self.lib.dyn_bam_stream_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def peek(path: str, max_records: int = 500000):
    """dyn_bam_peek: (set of aux tags, OR of the flags, records looked at) of the first records -- what
    bam.get_tags_from_bam / check_bam_* (bam.py:447-570) learn about a BAM."""
    lib = _lib.load()
    buf = C.create_string_buffer(2 * 4096)
    n_tags, flag_or, n_seen = C.c_int32(0), C.c_uint32(0), C.c_int64(0)
    _lib.check(lib.dyn_bam_peek(path.encode(), int(max_records), buf, len(buf), C.byref(n_tags), C.byref(flag_or), C.byref(n_seen)))
    raw = buf.raw[:2 * min(n_tags.value, 4096)]
    return {raw[i:i + 2].decode() for i in range(0, len(raw), 2)}, int(flag_or.value), int(n_seen.value)


def write_bam(path: str, refs, blob: np.ndarray, rec_off: np.ndarray, order=None, names=None, aux=None, mapq=None, flag=None,
              header_text: Optional[str] = None, level: int = 6, n_threads: int = 8, index: bool = True, raw=None, raw_start=None,
              raw_len=None, retag=None, gene_tag: Optional[str] = None, md=None):
    """dyn_bam_write(_ex): packed records (full layout, no mates) -> a BAM file (+ .bai).  refs: [(name, length)];
    names: list of str or None ("r<i>"); aux: list of bytes (pre-encoded tags) or a (bytes, uint32 offsets) pair;
    order: unit order in the file (coordinate order for a sorted BAM).  raw / raw_start / raw_len: units written from
    BAM alignment records as they are (raw_len[u] > 0); retag: (bytes, uint32 offsets) of "gene\0name\0" per unit whose
    gene tags are swapped first (swap_gene_tags, consensus.py:277-286)."""
    lib = _lib.load()
    n = len(rec_off) - 1
    if header_text is None:
        header_text = '@HD\tVN:1.4\tSO:coordinate\n' + ''.join(f'@SQ\tSN:{r}\tLN:{l}\n' for r, l in refs)
    ref_names = (C.c_char_p * max(len(refs), 1))(*[str(r).encode() for r, _ in refs])
    ref_lens = np.asarray([l for _, l in refs], dtype=np.int32)

    def pool(items):
        if items is None:
            return None, None
        if isinstance(items, tuple):
            return np.frombuffer(items[0], dtype=np.uint8), np.ascontiguousarray(items[1], dtype=np.uint32)
        enc = [x.encode() if isinstance(x, str) else bytes(x) for x in items]
        off = np.zeros(len(enc) + 1, dtype=np.uint32)
        np.cumsum([len(x) for x in enc], out=off[1:])
        return np.frombuffer(b''.join(enc) or b'\0', dtype=np.uint8), off

    n_data, n_off = pool(names)
    a_data, a_off = pool(aux)
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.uint32)
    order = None if order is None else np.ascontiguousarray(order, dtype=np.int64)
    mapq = None if mapq is None else np.ascontiguousarray(mapq, dtype=np.uint8)
    flag = None if flag is None else np.ascontiguousarray(flag, dtype=np.uint16)
    p = _lib.ptr
    o = _lib.BamWriteOpts(header_text=header_text.encode(), n_ref=len(refs), ref_names=C.cast(ref_names, C.POINTER(C.c_char_p)),
                          ref_lens=p(ref_lens), h_records=p(blob) if blob.size else None, h_rec_off=p(rec_off), n_units=n, h_order=p(order),
                          h_names=p(n_data), h_name_off=p(n_off), h_aux=p(a_data), h_aux_off=p(a_off), h_mapq=p(mapq), h_flag=p(flag),
                          level=int(level), n_threads=int(n_threads), write_index=int(bool(index)))
    keep = [ref_names, ref_lens, blob, rec_off, order, n_data, n_off, a_data, a_off, mapq, flag]
    if raw is not None:
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        raw_start = np.ascontiguousarray(raw_start, dtype=np.uint64)
        raw_len = np.ascontiguousarray(raw_len, dtype=np.uint32)
        o.h_raw, o.h_raw_start, o.h_raw_len = p(raw), p(raw_start), p(raw_len)
        keep += [raw, raw_start, raw_len]
    if retag is not None:
        r_data, r_off = pool(retag)
        o.h_retag, o.h_retag_off = p(r_data), p(r_off)
        o.gene_tag = gene_tag.encode() if gene_tag else None
        keep += [r_data, r_off]
    if md is not None:          # MD text per unit, written verbatim where not empty
        m_data, m_off = pool(md)
        o.h_md, o.h_md_off = p(m_data), p(m_off)
        keep += [m_data, m_off]
    _lib.check(lib.dyn_bam_write_ex(path.encode(), C.byref(o)))
    del keep
    return path


class _DeviceArray:
    """A library-owned device buffer seen through __cuda_array_interface__ (zero-copy torch.as_tensor)."""

    def __init__(self, pointer: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {'shape': (n,), 'typestr': typestr, 'data': (pointer, False), 'version': 2}


class DeviceBamStream:
    """dyn_bam_dev_*: the block range `part` of `n_parts` of a coordinate-sorted single-end BAM, inflated and packed on
    the GPU (csrc/bamdevice.cu).  Iterating yields dicts of torch tensors that VIEW library memory -- records (uint8),
    rec_off (int32 bit pattern of uint32, n + 1), barcode_key / umi_key (int64 bit patterns), gene_idx, score, ref_id,
    pos, hi (int32), flag (int16 bit pattern), gx_assigned (uint8) -- valid until the second-next chunk is requested, plus
    n_units / n_records_seen / compressed_bytes / inflated_bytes.  Raises DynastB200Error for files this path does not
    take (paired records, records that straddle BGZF blocks): use BamStream for those."""

    def __init__(self, ctx, path: str, barcode_tag: Optional[str] = None, umi_tag: Optional[str] = None,
                 gene_tag: Optional[str] = 'GX', require_gene: bool = False, part: int = 0, n_parts: int = 1,
                 chunk_bytes: int = 0, lean: bool = True, gene_ids: Optional[List[str]] = None, n_threads: int = 0):
        """n_threads: host threads that read the file into page-locked memory (0: the library's default, 16)."""
        self.ctx = ctx
        self.lib = ctx.lib
        enc = lambda s: s.encode() if s else None
        o = _lib.BamStreamOpts(barcode_tag=enc(barcode_tag), umi_tag=enc(umi_tag), gene_tag=enc(gene_tag),
                               require_gene=int(require_gene), n_threads=int(n_threads), part=int(part), n_parts=int(n_parts),
                               chunk_bytes=int(chunk_bytes), lean=int(lean), keys=1)
        self._gene_ids = None
        if gene_ids is not None:
            self._gene_ids = (C.c_char_p * max(len(gene_ids), 1))(*[g.encode() for g in gene_ids])
            o.gene_ids = C.cast(self._gene_ids, C.POINTER(C.c_char_p))
            o.n_gene_ids = len(gene_ids)
        h = C.c_void_p()
        _lib.check(self.lib.dyn_bam_dev_open(ctx.handle, path.encode(), C.byref(o), C.byref(h)))
        self.handle = h

    def _raw(self, field: int, dtype):
        nb = C.c_int64(0)
        p = self.lib.dyn_bam_dev_field(self.handle, field, C.byref(nb))
        if not p or nb.value == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nb.value,)).view(dtype).copy()

    @property
    def ref_len(self) -> np.ndarray:
        return self._raw(9, np.int32)

    @property
    def ref_name(self) -> np.ndarray:
        data, off = self._raw(28, np.uint8).tobytes(), self._raw(29, np.uint32)
        return np.array([data[off[i]:off[i + 1]].decode() for i in range(len(off) - 1)], dtype=object)

    def __iter__(self):
        import torch
        dev = torch.device('cuda', self.ctx.device)
        while True:
            ch = _lib.BamDevChunk()
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(self.lib.dyn_bam_dev_next(self.handle, C.byref(ch), stream))
            if ch.end == 2:
                return
            n = int(ch.n_units)
            view = lambda p, count, ts, dt: (torch.as_tensor(_DeviceArray(p, count, ts), device=dev) if count else torch.empty(0, dtype=dt, device=dev))
            out = dict(n_units=n, n_records_seen=int(ch.n_records_seen), compressed_bytes=int(ch.compressed_bytes),
                       inflated_bytes=int(ch.inflated_bytes))
            if n:
                out.update(
                    records=view(ch.d_records, int(ch.records_bytes), '|u1', torch.uint8), rec_off=view(ch.d_rec_off, n + 1, '<i4', torch.int32),
                    barcode_key=view(ch.d_barcode_key, n, '<i8', torch.int64), umi_key=view(ch.d_umi_key, n, '<i8', torch.int64),
                    gene_idx=view(ch.d_gene_idx, n, '<i4', torch.int32), score=view(ch.d_score, n, '<i4', torch.int32),
                    ref_id=view(ch.d_ref_id, n, '<i4', torch.int32), pos=view(ch.d_pos, n, '<i4', torch.int32),
                    hi=view(ch.d_hi, n, '<i4', torch.int32), flag=view(ch.d_flag, n, '<i2', torch.int16),
                    gx_assigned=view(ch.d_gx_assigned, n, '|u1', torch.uint8))
            yield out
            if ch.end:
                return

    def close(self):
        if self.handle:
            self.lib.dyn_bam_dev_close(self.handle)
            self.handle = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

"""Python face of the host BAM ingest (csrc/bampack.cpp, dyn_bam_pack): decode a coordinate-sorted BAM into the
packed records the kernels read plus the per-unit metadata the CSV outputs need."""
import ctypes as C
from typing import List, Optional

import numpy as np

from . import _lib

_FIELDS = {'blob': (0, np.uint8), 'rec_off': (1, np.uint32), 'ref_id': (2, np.int32), 'pos': (3, np.int32),
           'flag': (4, np.uint16), 'hi': (5, np.int32), 'score': (6, np.int32), 'gx_assigned': (7, np.uint8),
           'paired': (8, np.uint8), 'ref_len': (9, np.int32)}
_POOLS = {'read_id': (20, 21), 'barcode': (22, 23), 'umi': (24, 25), 'gene': (26, 27), 'ref_name': (28, 29)}


class PackedBam:
    """Result of dyn_bam_pack. Numeric fields are numpy views into library-owned memory (valid until close());
    string fields are decoded on demand."""

    def __init__(self, path: str, barcode_tag: Optional[str] = None, umi_tag: Optional[str] = None,
                 gene_tag: Optional[str] = 'GX', require_gene: bool = False, n_threads: int = 8):
        self.lib = _lib.load()
        h = C.c_void_p()
        enc = lambda s: s.encode() if s else None
        _lib.check(self.lib.dyn_bam_pack(path.encode(), enc(barcode_tag), enc(umi_tag), enc(gene_tag), int(require_gene),
                                         int(n_threads), C.byref(h)))
        self.handle = h
        self.n_units = int(self.lib.dyn_bam_n_units(h))
        self.n_records_seen = int(self.lib.dyn_bam_n_records_seen(h))
        self._cache = {}

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
        out = np.empty(len(off) - 1, dtype=object)
        for i in range(len(off) - 1):
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

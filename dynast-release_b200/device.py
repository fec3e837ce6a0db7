"""Thin Python wrappers over the C-ABI calls (device-pointer and host-buffer forms)."""
import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import Context, check, ptr

TALLY_NASC = 0x1
TALLY_EMIT_CONV = 0x2


def tally_reads_host(blob: np.ndarray, rec_off: np.ndarray, quality: int = 27, nasc: bool = False,
                     snps: Optional[np.ndarray] = None, emit_conv: bool = True, conv_capacity: Optional[int] = None,
                     device: int = 0):
    """dyn_tally_reads_host on numpy buffers. Returns dict(counts, conv_off, conv_n, conv_rec, conv_read, status)."""
    ctx = Context.get(device)
    n = len(rec_off) - 1
    blob = np.ascontiguousarray(blob, dtype=np.uint8)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.uint32)
    counts = np.zeros((n, 16), dtype=np.uint8)
    flags = (TALLY_NASC if nasc else 0) | (TALLY_EMIT_CONV if emit_conv else 0)
    cap = int(conv_capacity if conv_capacity is not None else max(16, len(blob) // 4))
    conv_off = np.zeros(n, dtype=np.uint32) if emit_conv else None
    conv_n = np.zeros(n, dtype=np.uint16) if emit_conv else None
    conv_rec = np.zeros(cap, dtype=np.uint64) if emit_conv else None
    conv_read = np.zeros(cap, dtype=np.uint32) if emit_conv else None
    total = C.c_uint64(0)
    status = C.c_uint32(0)
    snps = None if snps is None or len(snps) == 0 else np.ascontiguousarray(snps, dtype=np.uint64)
    check(
        ctx.lib.dyn_tally_reads_host(
            ctx.handle, ptr(blob), ptr(rec_off), n, ptr(snps), 0 if snps is None else len(snps), int(quality), flags,
            ptr(counts), ptr(conv_off), ptr(conv_n), ptr(conv_rec), ptr(conv_read), cap,
            C.cast(C.byref(total), C.c_void_p), C.cast(C.byref(status), C.c_void_p)
        )
    )
    t = int(total.value)
    out = dict(counts=counts, status=int(status.value), total=t)
    if emit_conv:
        out.update(conv_off=conv_off, conv_n=conv_n, conv_rec=conv_rec[:t], conv_read=conv_read[:t])
    return out


def em_pc_host(k: np.ndarray, n: np.ndarray, count: np.ndarray, off: np.ndarray, p_e: np.ndarray, nasc: bool = False,
               device: int = 0):
    """dyn_em_pc_host. Barcode b owns triplets [off[b], off[b+1]). Returns (p_c, status, iters)."""
    ctx = Context.get(device)
    B = len(off) - 1
    k = np.ascontiguousarray(k, dtype=np.int32)
    n = np.ascontiguousarray(n, dtype=np.int32)
    count = np.ascontiguousarray(count, dtype=np.int64)
    off = np.ascontiguousarray(off, dtype=np.int64)
    p_e = np.ascontiguousarray(p_e, dtype=np.float64)
    p_c = np.zeros(B, dtype=np.float64)
    status = np.zeros(B, dtype=np.int32)
    iters = np.zeros(B, dtype=np.int32)
    check(
        ctx.lib.dyn_em_pc_host(
            ctx.handle, ptr(k), ptr(n), ptr(count), ptr(off), B, ptr(p_e), 1 if nasc else 0, ptr(p_c), ptr(status),
            ptr(iters)
        )
    )
    return p_c, status, iters

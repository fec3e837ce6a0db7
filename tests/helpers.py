"""Shared helpers for the test-suite (golden fixture access, EM batches)."""
import csv
import gzip
import io
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, 'golden')
REF = os.path.join(GOLDEN, 'ref')


def ref_path(*parts):
    return os.path.join(REF, *parts)


def ref_text(*parts) -> str:
    """Text of a reference fixture (stored gzip-compressed when it is a CSV)."""
    p = ref_path(*parts)
    if os.path.exists(p + '.gz'):
        with gzip.open(p + '.gz', 'rt') as f:
            return f.read()
    with open(p) as f:
        return f.read()


def ref_rows(*parts):
    return list(csv.DictReader(io.StringIO(ref_text(*parts))))


def gene_infos():
    with open(os.path.join(GOLDEN, 'gene_infos.json')) as f:
        return json.load(f)


def em_golden():
    return np.load(os.path.join(GOLDEN, 'em_golden.npz'))


def synth_em_batch(n_barcodes, seed=2004, adversarial_every=100):
    """C4-like histograms: returns k, n, count, off, p_e."""
    rng = np.random.default_rng(seed)
    ks, ns, cs, off, pes = [], [], [], [0], []
    for b in range(n_barcodes):
        R = max(1000, int(rng.lognormal(np.log(3000), 0.5)))
        pi = rng.beta(2, 5)
        pcb = rng.uniform(0.02, 0.08)
        peb = rng.uniform(2e-4, 2e-3)
        n = np.clip(rng.poisson(23, R), 0, 150)
        lab = rng.random(R) < pi
        k = np.minimum(rng.binomial(n, np.where(lab, pcb, peb)), 30)
        if adversarial_every and b % adversarial_every == 0:
            k[0], n[0] = 30, 150
        key, cnt = np.unique(k.astype(np.int64) * 1024 + n, return_counts=True)
        ks.append((key // 1024).astype(np.int32))
        ns.append((key % 1024).astype(np.int32))
        cs.append(cnt.astype(np.int64))
        off.append(off[-1] + len(key))
        pes.append(peb)
    return (np.concatenate(ks), np.concatenate(ns), np.concatenate(cs), np.asarray(off, dtype=np.int64),
            np.asarray(pes, dtype=np.float64))


def oracle_em_batch(lib, k, n, count, off, p_e, nasc=False):
    B = len(off) - 1
    p_c = np.zeros(B)
    status = np.zeros(B, dtype=np.int32)
    k = np.ascontiguousarray(k, dtype=np.int32)
    n = np.ascontiguousarray(n, dtype=np.int32)
    count = np.ascontiguousarray(count, dtype=np.int64)
    off = np.ascontiguousarray(off, dtype=np.int64)
    p_e = np.ascontiguousarray(p_e, dtype=np.float64)
    lib.dyn_oracle_em_batch(k.ctypes.data, n.ctypes.data, count.ctypes.data, off.ctypes.data, B, p_e.ctypes.data,
                            1 if nasc else 0, p_c.ctypes.data, status.ctypes.data)
    return p_c, status


def materialize(tmp_path, *parts, name=None):
    """Write a (possibly gzip-stored) golden file to tmp_path uncompressed; returns the path."""
    src = os.path.join(GOLDEN, *parts)
    dst = os.path.join(str(tmp_path), name or parts[-1].replace('.gz', ''))
    if os.path.exists(src + '.gz'):
        src += '.gz'
    if src.endswith('.gz'):
        with gzip.open(src, 'rb') as f, open(dst, 'wb') as out:
            out.write(f.read())
    else:
        with open(src, 'rb') as f, open(dst, 'wb') as out:
            out.write(f.read())
    return dst


def golden_text(*parts) -> str:
    p = os.path.join(GOLDEN, *parts)
    if os.path.exists(p + '.gz'):
        p += '.gz'
    if p.endswith('.gz'):
        with gzip.open(p, 'rt') as f:
            return f.read()
    with open(p) as f:
        return f.read()


def synth_case_meta(name):
    import pickle
    with open(os.path.join(GOLDEN, 'synth_counts', name, 'meta.json')) as f:
        meta = json.load(f)
    meta['index'] = [tuple(x) for x in meta['index']]
    meta['snps'] = {c: {k: set(v) for k, v in d.items()} for c, d in meta['snps'].items()} or None
    return meta


def write_index(tmp_path, index, name='conversions.idx'):
    import pickle
    p = os.path.join(str(tmp_path), name)
    with gzip.open(p, 'wb') as f:
        pickle.dump(list(index), f)
    return p

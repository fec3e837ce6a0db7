"""Synthetic coordinate-sorted BAM for measuring the host ingest (dyn_bam_pack): the alignment records of a small BAM
are repeated `copies` times each (in place, so the file stays sorted) and written as BGZF
with zlib level 6, 64 KB of payload per block.
    python tools/make_bam.py <in.bam> <out.bam> <copies> [aligned]"""
import struct
import sys
import zlib

import numpy as np


def bgzf_blocks(data):
    """Inflated payloads of the BGZF blocks of `data`."""
    out, o = [], 0
    while o < len(data):
        xlen = struct.unpack_from('<H', data, o + 10)[0]
        bsize = None
        p = o + 12
        while p < o + 12 + xlen:
            si1, si2, slen = data[p], data[p + 1], struct.unpack_from('<H', data, p + 2)[0]
            if si1 == 66 and si2 == 67:
                bsize = struct.unpack_from('<H', data, p + 4)[0] + 1
            p += 4 + slen
        payload = data[o + 12 + xlen:o + bsize - 8]
        out.append(zlib.decompress(payload, -15))
        o += bsize
    return b''.join(out)


def bgzf_block(f, chunk, level=6):
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = c.compress(chunk) + c.flush()
    f.write(b'\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00' + struct.pack('<H', len(comp) + 25) + comp +
            struct.pack('<II', zlib.crc32(chunk) & 0xffffffff, len(chunk)))


def bgzf_write(f, payload, rec_len=None):
    """rec_len: all records of `payload` have this length -> blocks end on record boundaries, as htslib (and STAR)
    write them; None: plain 0xff00-byte blocks."""
    step = 0xff00 if not rec_len else max(rec_len, 0xff00 // rec_len * rec_len)
    for o in range(0, len(payload), step):
        bgzf_block(f, payload[o:o + step])


def main():
    src, dst, copies = sys.argv[1], sys.argv[2], int(sys.argv[3])
    aligned = len(sys.argv) > 4 and sys.argv[4] == 'aligned'      # blocks end on record boundaries
    raw = bgzf_blocks(open(src, 'rb').read())
    assert raw[:4] == b'BAM\x01'
    l_text = struct.unpack_from('<i', raw, 4)[0]
    o = 8 + l_text
    n_ref = struct.unpack_from('<i', raw, o)[0]
    o += 4
    for _ in range(n_ref):
        l_name = struct.unpack_from('<i', raw, o)[0]
        o += 4 + l_name + 4
    header, body = raw[:o], raw[o:]
    # records of the first reference only (a shifted copy must not leave its contig)
    recs, p = [], 0
    while p < len(body):
        bs = struct.unpack_from('<i', body, p)[0]
        recs.append(body[p:p + 4 + bs])
        p += 4 + bs
    by_ref = {}
    for r in recs:
        by_ref.setdefault(struct.unpack_from('<i', r, 4)[0], []).append(r)
    rng = np.random.default_rng(1)
    with open(dst, 'wb') as f:
        bgzf_write(f, header)
        n = 0
        for ref in sorted(k for k in by_ref if k >= 0):
            buf = bytearray()
            for r in by_ref[ref]:                 # every record `copies` times in a row: same position, still sorted
                block = np.frombuffer(r * copies, dtype=np.uint8).reshape(copies, len(r)).copy()
                l_name, n_cig, l_seq = r[12], struct.unpack_from('<H', r, 16)[0], struct.unpack_from('<i', r, 20)[0]
                q0 = 36 + l_name + 4 * n_cig + (l_seq + 1) // 2
                # fresh base qualities per copy, drawn like a sequencer's (mostly 37-40): keeps the stream realistically
                # compressible instead of 8 000 identical records
                block[:, q0:q0 + l_seq] = np.where(rng.random((copies, l_seq)) < 0.85, rng.integers(36, 41, (copies, l_seq)),
                                                   rng.integers(2, 36, (copies, l_seq)))
                n += copies
                if aligned:
                    bgzf_write(f, block.tobytes(), rec_len=len(r))
                    continue
                buf += block.tobytes()
                if len(buf) > (64 << 20):
                    bgzf_write(f, bytes(buf))
                    buf = bytearray()
            bgzf_write(f, bytes(buf))
        f.write(bytes.fromhex('1f8b08040000000000ff0600424302001b0003000000000000000000'))
    print(n, 'records')


if __name__ == '__main__':
    main()

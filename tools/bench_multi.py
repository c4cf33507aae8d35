#!/usr/bin/env python3
"""Multi-stream archives through the Swift-mirror API (host memory in, host memory out): the API-level form of BASELINE configs 4
and 5 — `BZip2.multiDecompress` on N concatenated single-block 900 KB streams and `XZArchive.splitUnarchive` on N concatenated
1 MiB-dictionary streams.  Streams are discovered up front (signature scan / stream indexes) and decoded as one batch."""
import bz2
import json
import lzma
import os
import sys
import time
from multiprocessing import Pool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def _bz(seed):
    import helpers as H
    raw = H.textlike(900000, seed)
    return bz2.compress(raw, 9), raw


def _xz(seed):
    import helpers as H
    raw = H.textlike(1 << 20, seed)
    return lzma.compress(raw, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC64,
                         filters=[{"id": lzma.FILTER_LZMA2, "preset": 6, "dict_size": 1 << 20}]), raw


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    distinct = 16
    with Pool(min(os.cpu_count() or 1, 16)) as pool:
        bz = pool.map(_bz, range(4, 4 + distinct))
        xz = pool.map(_xz, range(5, 5 + distinct))
    import swco
    from swcompression_b200 import BZip2, XZArchive
    for name, units, fn, ofn in (("BZip2.multiDecompress, %d x 900 KB single-block streams" % n, bz, BZip2.multiDecompress, swco.bzip2_decompress),
                                 ("XZArchive.splitUnarchive, %d x 1 MiB streams (CRC64, dict 1 MiB)" % n, xz, XZArchive.splitUnarchive, swco.xz_unarchive)):
        data = b"".join(units[i % distinct][0] for i in range(n))
        fn(data[:len(units[0][0]) + len(units[1][0])])           # warm-up on two streams
        t = time.perf_counter()
        parts = fn(data)
        dt = time.perf_counter() - t
        assert len(parts) == n and all(parts[i] == units[i % distinct][1] for i in range(0, n, 37))
        out_bytes = sum(len(p) for p in parts)
        t = time.perf_counter()
        ost, oout, _ = ofn(units[0][0])
        cdt = time.perf_counter() - t
        assert ost == 0
        print(json.dumps({"case": name, "streams": n, "compressed_bytes": len(data), "decompressed_bytes": out_bytes, "wall_ms": dt * 1e3,
                          "decompressed_MBps": out_bytes / dt / 1e6, "cpu_port_1thread_MBps": len(oout) / cdt / 1e6}))


if __name__ == "__main__":
    main()

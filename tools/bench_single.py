#!/usr/bin/env python3
"""Single-stream latency through the Swift-mirror API (the shape of the reference's own `swcomp benchmark run un-gzip / un-bz2 /
un-xz / lz4` rows: ONE archive, host memory in, host memory out).  Prints MB/s of decompressed output and compressed input."""
import bz2
import gzip
import json
import lzma
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import helpers as H  # noqa: E402


def best(fn, arg, reps=3):
    fn(arg)
    ts = []
    out = None
    for _ in range(reps):
        t = time.perf_counter()
        out = fn(arg)
        ts.append(time.perf_counter() - t)
    return min(ts), out


def main():
    import swco
    from swcompression_b200 import BZip2, GzipArchive, LZ4, XZArchive
    mib = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    raw = b"".join(H.textlike(1 << 20, 7000 + i) for i in range(mib))
    blocks = [raw[i:i + (4 << 20)] for i in range(0, len(raw), 4 << 20)]
    cases = [
        ("gzip (Deflate, 1 member)", gzip.compress(raw, 6), GzipArchive.unarchive, swco.gzip_unarchive),
        ("bzip2 -9 (%d blocks)" % ((len(raw) + 899999) // 900000), bz2.compress(raw, 9), BZip2.decompress, swco.bzip2_decompress),
        ("xz -6 (1 block)", lzma.compress(raw, preset=6), XZArchive.unarchive, swco.xz_unarchive),
        ("lz4 frame, independent 4 MiB blocks", H.lz4_frame_independent(blocks, bd=0x70), LZ4.decompress, swco.lz4_decompress),
    ]
    for name, comp, fn, ofn in cases:
        dt, out = best(fn, comp)
        assert out == raw, name
        t = time.perf_counter()
        ost, oout, _ = ofn(comp)
        cdt = time.perf_counter() - t
        assert ost == 0 and oout == raw
        print(json.dumps({"case": name, "raw_MiB": mib, "compressed_bytes": len(comp), "gpu_ms": dt * 1e3,
                          "gpu_out_MBps": len(raw) / dt / 1e6, "gpu_in_MBps": len(comp) / dt / 1e6,
                          "cpu_port_1thread_out_MBps": len(raw) / cdt / 1e6}))


if __name__ == "__main__":
    main()

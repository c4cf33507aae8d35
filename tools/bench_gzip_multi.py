#!/usr/bin/env python3
"""BASELINE config 2 through the public API: a BGZF-shaped archive of N gzip members of 64 KiB each, host memory in, host memory
out, `GzipArchive.multiUnarchive`.  Members are found by signature, decoded as one batch and validated in order."""
import gzip
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    import helpers as H
    import swco
    from swcompression_b200 import GzipArchive
    distinct = 256
    raws = [H.textlike(65536, 3000 + i) for i in range(distinct)]
    blobs = [gzip.compress(r, 6) for r in raws]
    data = b"".join(blobs[i % distinct] for i in range(n))
    GzipArchive.multiUnarchive(data[:sum(len(b) for b in blobs[:64])])       # warm-up
    # (a) the C ABI call alone (what the Swift binding would pay), (b) the Python mirror (adds bytes copies in and out)
    import ctypes as C
    from swcompression_b200 import _lib
    L = _lib.lib()
    buf, nbytes = _lib.inbuf(data)
    abi_ms = []
    for rep in range(2):
        out, out_len, ends, cnt = C.c_void_p(), C.c_size_t(0), C.c_void_p(), C.c_size_t(0)
        t = time.perf_counter()
        st = L.swc_gzip_multi_unarchive(buf, nbytes, C.byref(out), C.byref(out_len), C.byref(ends), C.byref(cnt))
        abi_ms.append((time.perf_counter() - t) * 1e3)
        assert st == 0 and cnt.value == n
        abi_bytes = out_len.value
        L.swc_free(out); L.swc_free(ends)
    del buf
    t = time.perf_counter()
    parts = GzipArchive.multiUnarchive(data)
    dt = time.perf_counter() - t
    assert len(parts) == n and all(parts[i].data == raws[i % distinct] for i in range(0, n, 101))
    out_bytes = sum(len(p.data) for p in parts)
    sample = b"".join(blobs[:64])
    t = time.perf_counter()
    ost, _, whole = swco.gzip_multi_unarchive(sample)
    cdt = time.perf_counter() - t
    assert ost == 0
    print(json.dumps({"case": "GzipArchive.multiUnarchive, %d x 64 KiB members (BGZF-shaped)" % n, "members": n,
                      "compressed_bytes": len(data), "decompressed_bytes": out_bytes, "c_abi_ms": min(abi_ms),
                      "c_abi_decompressed_MBps": abi_bytes / min(abi_ms) / 1e3, "python_mirror_ms": dt * 1e3, "cpu_port_1thread_MBps": len(whole) / cdt / 1e6}))


if __name__ == "__main__":
    main()

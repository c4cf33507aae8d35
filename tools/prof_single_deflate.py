import ctypes as C, sys, time, gzip
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import helpers as H
from swcompression_b200 import _lib
L = _lib.lib()
L.swc_timing_collect.argtypes = [C.c_void_p, C.c_int32]
raw = b"".join(H.textlike(1 << 20, 7000 + i) for i in range(16))
comp = H.raw_deflate(raw)
for rep in range(3):
    t0 = time.perf_counter()
    buf, n = _lib.inbuf(comp)
    t1 = time.perf_counter()
    out, out_len, used = C.c_void_p(), C.c_size_t(0), C.c_size_t(0)
    L.swc_timing_enable(1)
    st = L.swc_deflate_decompress(buf, n, 0, C.byref(out), C.byref(out_len), C.byref(used))
    t2 = time.perf_counter()
    tb = (C.c_float * 16)(); k = L.swc_timing_collect(tb, 16); L.swc_timing_enable(0)
    res = _lib.take(out, out_len)
    t3 = time.perf_counter()
    assert st == 0 and res == raw
    print("rep", rep, "inbuf %.1f  C call %.1f  take %.1f ms; kernels" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), [round(tb[i], 1) for i in range(k)])

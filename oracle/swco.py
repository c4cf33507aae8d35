"""ctypes binding of the CPU oracle (oracle/libswco.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Every function returns (status, output_bytes, extra) where status is an include/swc_status.h code.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Buf(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("cap", C.c_size_t)]


def build(force=False):
    so = os.path.join(_HERE, "libswco.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libswco.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.swco_crc32.restype = C.c_uint32
        _LIB.swco_bzip2_crc32.restype = C.c_uint32
        _LIB.swco_crc64.restype = C.c_uint64
        _LIB.swco_adler32.restype = C.c_uint32
        _LIB.swco_xxh32.restype = C.c_uint32
        _LIB.swco_sha256.restype = None
    return _LIB


def _take(buf):
    out = C.string_at(buf.data, buf.len) if buf.len else b""
    C.CDLL(None).free(C.c_void_p(buf.data))
    return out


def _ptr(b):
    return (C.c_uint8 * max(len(b), 1)).from_buffer_copy(bytes(b) + (b"\0" if not b else b""))


def deflate_decompress(data, start_bit=0):
    buf, used = _Buf(), C.c_uint64(0)
    st = lib().swco_deflate_decompress(_ptr(data), C.c_size_t(len(data)), C.c_uint64(start_bit), C.byref(buf), C.byref(used))
    return st, _take(buf), used.value


def bzip2_decompress(data, start_bit=0):
    buf, used = _Buf(), C.c_uint64(0)
    st = lib().swco_bzip2_decompress(_ptr(data), C.c_size_t(len(data)), C.c_uint64(start_bit), C.byref(buf), C.byref(used))
    return st, _take(buf), used.value


def _multi(fn, data, *pre):
    buf, n = _Buf(), C.c_size_t(0)
    ends = (C.c_size_t * 65536)()
    st = fn(_ptr(data), C.c_size_t(len(data)), *pre, C.byref(buf), ends, C.c_size_t(65536), C.byref(n))
    whole = _take(buf)
    parts, prev = [], 0
    for i in range(n.value):
        parts.append(whole[prev:ends[i]])
        prev = ends[i]
    return st, parts, whole


def bzip2_multi_decompress(data):
    return _multi(lib().swco_bzip2_multi_decompress, data)


def lz4_block(data, dictionary=None):
    buf = _Buf()
    d = dictionary or b""
    st = lib().swco_lz4_block(_ptr(data), C.c_size_t(len(data)), _ptr(d), C.c_size_t(len(d)), C.byref(buf))
    return st, _take(buf), None


def _dict_args(dictionary, dictionary_id):
    if dictionary is None:
        dp, dl = None, 0
    else:
        dp, dl = _ptr(dictionary), len(dictionary)
    return dp, C.c_size_t(dl), C.c_int(0 if dictionary_id is None else 1), C.c_uint32(dictionary_id or 0)


def lz4_decompress(data, dictionary=None, dictionary_id=None):
    buf, used = _Buf(), C.c_size_t(0)
    st = lib().swco_lz4_decompress(_ptr(data), C.c_size_t(len(data)), *_dict_args(dictionary, dictionary_id), C.byref(buf), C.byref(used))
    return st, _take(buf), used.value


def lz4_multi_decompress(data, dictionary=None, dictionary_id=None):
    return _multi(lib().swco_lz4_multi_decompress, data, *_dict_args(dictionary, dictionary_id))


def lzma_decompress(data):
    buf, used = _Buf(), C.c_size_t(0)
    st = lib().swco_lzma_decompress(_ptr(data), C.c_size_t(len(data)), C.byref(buf), C.byref(used))
    return st, _take(buf), used.value


def lzma_decompress_raw(data, lc, lp, pb, dict_size, uncompressed_size=None):
    buf, used = _Buf(), C.c_size_t(0)
    us = -1 if uncompressed_size is None else uncompressed_size
    st = lib().swco_lzma_decompress_raw(_ptr(data), C.c_size_t(len(data)), lc, lp, pb, C.c_int64(dict_size), C.c_int64(us), C.byref(buf), C.byref(used))
    return st, _take(buf), used.value


def lzma2_decompress(data):
    buf, used = _Buf(), C.c_size_t(0)
    st = lib().swco_lzma2_decompress(_ptr(data), C.c_size_t(len(data)), C.byref(buf), C.byref(used))
    return st, _take(buf), used.value


def lzma2_decompress_raw(data, dict_byte):
    buf, used = _Buf(), C.c_size_t(0)
    st = lib().swco_lzma2_decompress_raw(_ptr(data), C.c_size_t(len(data)), C.c_uint8(dict_byte), C.byref(buf), C.byref(used))
    return st, _take(buf), used.value


def gzip_unarchive(data):
    buf, used = _Buf(), C.c_size_t(0)
    st = lib().swco_gzip_unarchive(_ptr(data), C.c_size_t(len(data)), C.byref(buf), C.byref(used))
    return st, _take(buf), used.value


def gzip_multi_unarchive(data):
    return _multi(lib().swco_gzip_multi_unarchive, data)


def zlib_unarchive(data):
    buf = _Buf()
    st = lib().swco_zlib_unarchive(_ptr(data), C.c_size_t(len(data)), C.byref(buf))
    return st, _take(buf), None


def xz_unarchive(data):
    buf = _Buf()
    st = lib().swco_xz_unarchive(_ptr(data), C.c_size_t(len(data)), C.byref(buf))
    return st, _take(buf), None


def xz_split_unarchive(data):
    return _multi(lib().swco_xz_split_unarchive, data)


def crc32(data, prev=0):
    return lib().swco_crc32(_ptr(data), C.c_size_t(len(data)), C.c_uint32(prev))


def bzip2_crc32(data):
    return lib().swco_bzip2_crc32(_ptr(data), C.c_size_t(len(data)))


def crc64(data):
    return lib().swco_crc64(_ptr(data), C.c_size_t(len(data)))


def adler32(data):
    return lib().swco_adler32(_ptr(data), C.c_size_t(len(data)))


def xxh32(data):
    return lib().swco_xxh32(_ptr(data), C.c_size_t(len(data)))


def sha256(data):
    dg = (C.c_uint8 * 32)()
    lib().swco_sha256(_ptr(data), C.c_size_t(len(data)), dg)
    return bytes(dg)


CODECS = {"deflate": 0, "lz4_block": 1, "bzip2": 2, "lzma2": 3, "xz": 4, "gzip": 5}


def batch_mt(codec, units, total, threads, aux=0):
    """Decode `total` units (wrapping over `units`) on `threads` pthreads inside C (oracle/batch_mt.c): no interpreter in the
    timed loop.  -> (seconds, decoded bytes, failures)."""
    import numpy as np
    lens = np.fromiter((len(u) for u in units), dtype=np.uint64, count=len(units))
    offs = np.zeros(len(units), dtype=np.uint64)
    if len(units) > 1:
        offs[1:] = np.cumsum(lens[:-1])
    blob = np.frombuffer(b"".join(units) + b"\0" * 16, dtype=np.uint8)
    sec, nbytes, fails = C.c_double(0), C.c_uint64(0), C.c_uint64(0)
    rc = lib().swco_batch_mt(C.c_int(CODECS[codec]), blob.ctypes.data_as(C.c_void_p), offs.ctypes.data_as(C.c_void_p),
                             lens.ctypes.data_as(C.c_void_p), C.c_uint64(len(units)), C.c_uint64(total), C.c_int(aux),
                             C.c_int(threads), C.byref(sec), C.byref(nbytes), C.byref(fails))
    if rc != 0:
        raise RuntimeError("swco_batch_mt could not start its threads")
    return sec.value, nbytes.value, fails.value


class _ZipEntry(C.Structure):
    _fields_ = [("name_off", C.c_uint64), ("name_len", C.c_uint64), ("comment_off", C.c_uint64), ("comment_len", C.c_uint64),
                ("data_off", C.c_uint64), ("data_len", C.c_uint64), ("size", C.c_uint64), ("crc", C.c_uint32),
                ("external_attrs", C.c_uint32), ("method", C.c_uint16), ("version_made_by", C.c_uint16),
                ("internal_attrs", C.c_uint16), ("dos_time", C.c_uint16), ("dos_date", C.c_uint16), ("is_directory", C.c_uint8),
                ("utf8", C.c_uint8)]


def zip_open(data, info_only=False, max_entries=1 << 16):
    """ZipContainer.open(container:) / info(container:) -> (status, [dict(name, comment, data, size, crc, method, is_directory, utf8)])"""
    data = bytes(data)
    buf, n = _Buf(), C.c_size_t(0)
    ents = (_ZipEntry * max_entries)()
    st = lib().swco_zip_open(_ptr(data), C.c_size_t(len(data)), C.byref(buf), ents, C.c_size_t(max_entries), C.byref(n), C.c_int(1 if info_only else 0))
    whole = _take(buf)
    out = []
    for i in range(min(n.value, max_entries)):
        e = ents[i]
        out.append(dict(name=data[e.name_off:e.name_off + e.name_len], comment=data[e.comment_off:e.comment_off + e.comment_len],
                        data=None if (e.is_directory or info_only) else whole[e.data_off:e.data_off + e.data_len],
                        size=e.size, crc=e.crc, method=e.method, is_directory=bool(e.is_directory), utf8=bool(e.utf8),
                        external_attrs=e.external_attrs, version_made_by=e.version_made_by, dos_time=e.dos_time, dos_date=e.dos_date))
    return st, out

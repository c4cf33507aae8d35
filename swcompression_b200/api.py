"""Host-side mirror of the reference's Swift API for the decode hot path (same type and method names, same argument
meaning, same error cases), routed through the C ABI of libswcgpu.so.

    Swift (reference)                                   here
    Deflate.decompress(data:)                           Deflate.decompress(data)
    BZip2.decompress(data:) / multiDecompress(data:)    BZip2.decompress(data) / BZip2.multiDecompress(data)
    LZMA.decompress(data:[properties:uncompressedSize:]) LZMA.decompress(data[, properties, uncompressedSize])
    LZMA2.decompress(data:)                             LZMA2.decompress(data)
    LZ4.decompress(data:[dictionary:dictionaryID:])     LZ4.decompress(data[, dictionary, dictionaryID])
    LZ4.multiDecompress(data:dictionary:dictionaryID:)  LZ4.multiDecompress(...)
    GzipArchive.unarchive / multiUnarchive              GzipArchive.unarchive / multiUnarchive
    ZlibArchive.unarchive                               ZlibArchive.unarchive
    XZArchive.unarchive / splitUnarchive                XZArchive.unarchive / splitUnarchive
"""
import ctypes as C
from dataclasses import dataclass

from . import _lib
from .errors import check, error_for

_PAYLOAD_CODES = {210, 503, 605, 705, 807}


def _single(fn, data, *extra, consumed=True):
    L = _lib.lib()
    buf, n = _lib.inbuf(data)
    out, out_len, used = C.c_void_p(), C.c_size_t(0), C.c_size_t(0)
    args = [buf, n, *extra, C.byref(out), C.byref(out_len)] + ([C.byref(used)] if consumed else [])
    st = getattr(L, fn)(*args)
    payload = _lib.take(out, out_len)
    if st != 0:
        raise error_for(st, payload if st in _PAYLOAD_CODES else None)
    return payload, used.value


def _multi(fn, data, *extra):
    L = _lib.lib()
    buf, n = _lib.inbuf(data)
    out, out_len, ends, cnt = C.c_void_p(), C.c_size_t(0), C.c_void_p(), C.c_size_t(0)
    st = getattr(L, fn)(buf, n, *extra, C.byref(out), C.byref(out_len), C.byref(ends), C.byref(cnt))
    whole = _lib.take(out, out_len)
    e = _lib.take_sizes(ends, cnt)
    parts, prev = [], 0
    for x in e:
        parts.append(whole[prev:x])
        prev = x
    if st != 0:
        raise error_for(st, parts if st in _PAYLOAD_CODES else None)
    return parts


class Deflate:
    """Sources/Deflate/Deflate.swift:10-28"""

    @staticmethod
    def decompress(data):
        return _single("swc_deflate_decompress", data, C.c_size_t(0))[0]

    @staticmethod
    def decompress_from(data, start_bit):
        """reader-based overload (Deflate.swift:30): returns (bytes, consumed_bits)"""
        return _single("swc_deflate_decompress", data, C.c_size_t(start_bit))


class BZip2:
    """Sources/BZip2/BZip2.swift:10-48"""

    @staticmethod
    def decompress(data):
        return _single("swc_bzip2_decompress", data, C.c_size_t(0))[0]

    @staticmethod
    def multiDecompress(data):
        return _multi("swc_bzip2_multi_decompress", data)


@dataclass
class LZMAProperties:
    """Sources/LZMA/LZMAProperties.swift:9-47"""
    lc: int = 3
    lp: int = 0
    pb: int = 2
    dictionarySize: int = 1 << 24

    def __post_init__(self):
        if self.dictionarySize < (1 << 12):     # the reference's didSet clamp (it does not fire in Swift's init, but the
            pass                                # memberwise init assigns through the property: LZMAProperties.swift:39-44)


class LZMA:
    """Sources/LZMA/LZMA.swift:10-73"""

    @staticmethod
    def decompress(data, properties=None, uncompressedSize=None):
        if properties is None:
            return _single("swc_lzma_decompress", data)[0]
        us = -1 if uncompressedSize is None else int(uncompressedSize)
        return _single("swc_lzma_decompress_raw", data, C.c_int32(properties.lc), C.c_int32(properties.lp), C.c_int32(properties.pb),
                       C.c_int64(properties.dictionarySize), C.c_int64(us))[0]


class LZMA2:
    """Sources/LZMA2/LZMA2.swift:10-36"""

    @staticmethod
    def decompress(data):
        return _single("swc_lzma2_decompress", data)[0]


def _dict_args(dictionary, dictionaryID):
    if dictionary is None:
        return [None, C.c_size_t(0), C.c_int32(0 if dictionaryID is None else 1), C.c_uint32(dictionaryID or 0)], None
    buf, n = _lib.inbuf(dictionary)
    return [buf, C.c_size_t(n), C.c_int32(0 if dictionaryID is None else 1), C.c_uint32(dictionaryID or 0)], buf


class LZ4:
    """Sources/LZ4/LZ4.swift:33-146"""

    @staticmethod
    def decompress(data, dictionary=None, dictionaryID=None):
        args, _keep = _dict_args(dictionary, dictionaryID)
        return _single("swc_lz4_decompress", data, *args)[0]

    @staticmethod
    def multiDecompress(data, dictionary=None, dictionaryID=None):
        args, _keep = _dict_args(dictionary, dictionaryID)
        return _multi("swc_lz4_multi_decompress", data, *args)


class GzipArchive:
    """Sources/GZip/GzipArchive.swift:10-77 (members are returned as their data; header metadata is not on the hot path)"""

    @staticmethod
    def unarchive(archive):
        return _single("swc_gzip_unarchive", archive)[0]

    @staticmethod
    def multiUnarchive(archive):
        return _multi("swc_gzip_multi_unarchive", archive)


class ZlibArchive:
    """Sources/Zlib/ZlibArchive.swift:10-42"""

    @staticmethod
    def unarchive(archive):
        return _single("swc_zlib_unarchive", archive, consumed=False)[0]


class XZArchive:
    """Sources/XZ/XZArchive.swift:10-88"""

    @staticmethod
    def unarchive(archive):
        return _single("swc_xz_unarchive", archive, consumed=False)[0]

    @staticmethod
    def splitUnarchive(archive):
        return _multi("swc_xz_split_unarchive", archive)


# ---- checks (CheckSums.swift / XxHash32.swift / Sha256.swift) ----
def _check32(fn, data):
    buf, n = _lib.inbuf(data)
    v = C.c_uint32(0)
    check(getattr(_lib.lib(), fn)(buf, n, C.byref(v)))
    return v.value


def crc32(data):
    return _check32("swc_crc32", data)


def bzip2_crc32(data):
    return _check32("swc_bzip2_crc32", data)


def adler32(data):
    return _check32("swc_adler32", data)


def xxh32(data):
    return _check32("swc_xxh32", data)


def crc64(data):
    buf, n = _lib.inbuf(data)
    v = C.c_uint64(0)
    check(_lib.lib().swc_crc64(buf, n, C.byref(v)))
    return v.value


def sha256(data):
    buf, n = _lib.inbuf(data)
    dg = (C.c_uint8 * 32)()
    check(_lib.lib().swc_sha256(buf, n, dg))
    return bytes(dg)

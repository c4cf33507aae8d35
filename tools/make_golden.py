#!/usr/bin/env python3
"""Regenerate the reference's Git-LFS test fixtures that can be reproduced offline and commit them as golden vectors.

The reference's `Tests/Test Files/**` are LFS *pointer* stubs (sha256 oid + size).  For the fixtures below the exact
bytes can be rebuilt with the system zlib / bzip2 / xz / liblz4 (recipes: SURVEY.md Appendix B); each rebuilt file is
accepted only if its sha256 equals the pointer's oid, so these are the reference's true fixtures, with known answers
(`Tests/Constants.swift:10-20`).  Output: tests/golden/<Fmt>/<name> + tests/golden/manifest.json.

Run here (needs /root/reference for the pointers); the GPU box only reads the committed files.
"""
import ctypes as C
import hashlib
import json
import lzma
import os
import struct
import subprocess
import sys
import zlib

REF = "/root/reference/Tests/Test Files"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ANSWERS = {
    "test1": b"Hello, World!\n",
    "test5": b"",
    "test6": bytes(1 << 20),
}


def pointer(rel):
    with open(os.path.join(REF, rel), "rb") as f:
        txt = f.read().decode()
    oid = size = None
    for line in txt.splitlines():
        if line.startswith("oid sha256:"):
            oid = line.split(":", 1)[1].strip()
        if line.startswith("size "):
            size = int(line.split()[1])
    return oid, size


def raw_deflate(data, level=6):
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    return c.compress(data) + c.flush()


def gz(name, mtime, payload_raw):
    body = raw_deflate(payload_raw)
    return (b"\x1f\x8b\x08\x08" + struct.pack("<I", mtime) + b"\x00\x03" + name.encode() + b"\x00" + body +
            struct.pack("<II", zlib.crc32(payload_raw), len(payload_raw) & 0xFFFFFFFF))


def tool(cmd, data):
    return subprocess.run(cmd, input=data, stdout=subprocess.PIPE, check=True).stdout


# ---- liblz4 frame API via ctypes (LZ4F_preferences_t layout of lz4 1.9.x) ----
class FrameInfo(C.Structure):
    _fields_ = [("blockSizeID", C.c_int), ("blockMode", C.c_int), ("contentChecksumFlag", C.c_int), ("frameType", C.c_int),
                ("contentSize", C.c_ulonglong), ("dictID", C.c_uint), ("blockChecksumFlag", C.c_int)]


class Prefs(C.Structure):
    _fields_ = [("frameInfo", FrameInfo), ("compressionLevel", C.c_int), ("autoFlush", C.c_uint),
                ("favorDecSpeed", C.c_uint), ("reserved", C.c_uint * 3)]


_lz4 = None


def lz4():
    global _lz4
    if _lz4 is None:
        _lz4 = C.CDLL("liblz4.so.1")
        _lz4.LZ4F_compressFrameBound.restype = C.c_size_t
        _lz4.LZ4F_compressFrameBound.argtypes = [C.c_size_t, C.c_void_p]
        _lz4.LZ4F_compressFrame.restype = C.c_size_t
        _lz4.LZ4F_compressFrame.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        _lz4.LZ4F_compressBound.restype = C.c_size_t
        _lz4.LZ4F_compressBound.argtypes = [C.c_size_t, C.c_void_p]
        _lz4.LZ4F_createCompressionContext.argtypes = [C.POINTER(C.c_void_p), C.c_uint]
        _lz4.LZ4F_compressBegin.restype = C.c_size_t
        _lz4.LZ4F_compressBegin.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _lz4.LZ4F_compressUpdate.restype = C.c_size_t
        _lz4.LZ4F_compressUpdate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        _lz4.LZ4F_compressEnd.restype = C.c_size_t
        _lz4.LZ4F_compressEnd.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        _lz4.LZ4F_freeCompressionContext.argtypes = [C.c_void_p]
        _lz4.LZ4_compressBound.restype = C.c_int
        _lz4.LZ4_compress_default.restype = C.c_int
        _lz4.LZ4_compress_default.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_int]
    return _lz4


def lz4_frame(data, bsid, linked, content_ck, content_size, block_ck):
    p = Prefs()
    p.frameInfo.blockSizeID = bsid
    p.frameInfo.blockMode = 0 if linked else 1
    p.frameInfo.contentChecksumFlag = content_ck
    p.frameInfo.contentSize = len(data) if content_size else 0
    p.frameInfo.blockChecksumFlag = block_ck
    L = lz4()
    cap = L.LZ4F_compressFrameBound(len(data), C.byref(p))
    dst = C.create_string_buffer(cap)
    n = L.LZ4F_compressFrame(dst, cap, data, len(data), C.byref(p))
    return dst.raw[:n]


def lz4_frame_streaming(data, bsid, chunk):
    p = Prefs()
    p.frameInfo.blockSizeID = bsid
    p.frameInfo.blockMode = 0
    p.frameInfo.contentSize = len(data)
    L = lz4()
    ctx = C.c_void_p()
    assert L.LZ4F_createCompressionContext(C.byref(ctx), 100) == 0
    cap = L.LZ4F_compressBound(chunk, C.byref(p)) + 64
    dst = C.create_string_buffer(cap)
    out = bytearray()
    n = L.LZ4F_compressBegin(ctx, dst, cap, C.byref(p)); out += dst.raw[:n]
    for i in range(0, len(data), chunk):
        piece = data[i:i + chunk]
        n = L.LZ4F_compressUpdate(ctx, dst, cap, piece, len(piece), None); out += dst.raw[:n]
    n = L.LZ4F_compressEnd(ctx, dst, cap, None); out += dst.raw[:n]
    L.LZ4F_freeCompressionContext(ctx)
    return bytes(out)


def lz4_block(data):
    L = lz4()
    cap = L.LZ4_compressBound(len(data))
    dst = C.create_string_buffer(cap)
    n = L.LZ4_compress_default(data, dst, len(data), cap)
    assert n > 0
    return dst.raw[:n]


def lz4_legacy(data):
    out = bytearray(b"\x02\x21\x4c\x18")
    for i in range(0, len(data), 8 << 20):
        b = lz4_block(data[i:i + (8 << 20)])
        out += struct.pack("<I", len(b)) + b
    return bytes(out)


def recipes():
    t1, t5, t6 = ANSWERS["test1"], ANSWERS["test5"], ANSWERS["test6"]
    z5m = bytes(5 * 1024 * 1024)
    r = {}
    r["Deflate/test6.deflate"] = (lambda: raw_deflate(t6), "test6")
    r["GZip/test1.gz"] = (lambda: gz("test1.answer", 1482698300, t1), "test1")
    r["GZip/test5.gz"] = (lambda: gz("test5.answer", 1482698242, t5), "test5")
    r["GZip/test6.gz"] = (lambda: gz("test6.answer", 1511554495, t6), "test6")
    r["GZip/minimal.gz"] = (lambda: b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x04\x03\x03\x00" + bytes(8), "test5")
    r["Zlib/test_empty.zlib"] = (lambda: zlib.compress(b"", 9), "test5")
    r["Zlib/test.zlib"] = (lambda: b"\x78\x9c", None)
    for k, a in (("test1", t1), ("test5", t5), ("test6", t6)):
        r[f"BZip2/{k}.bz2"] = (lambda a=a: tool(["bzip2", "-9", "-c"], a), k)
        r[f"XZ/{k}.xz"] = (lambda a=a: tool(["xz", "-c"], a), k)
    r["LZMA/test_empty.lzma"] = (lambda: tool(["xz", "--format=lzma", "-c"], b""), "test5")
    r["LZ4/test1.lz4"] = (lambda: lz4_frame(t1, 4, True, 1, True, 1), "test1")
    r["LZ4/test5.lz4"] = (lambda: lz4_frame(t5, 4, True, 1, False, 1), "test5")
    r["LZ4/test6.lz4"] = (lambda: lz4_frame(t6, 6, True, 1, True, 1), "test6")
    for b in (4, 5, 6, 7):
        r[f"LZ4/test_B{b}.lz4"] = (lambda b=b: lz4_frame(z5m, b, False, 0, True, 0), "zeros5m")
    for b, chunk in ((4, 64 << 10), (5, 256 << 10), (6, 1 << 20)):
        r[f"LZ4/test_B{b}_BD.lz4"] = (lambda b=b, chunk=chunk: lz4_frame_streaming(z5m, b, chunk), "zeros5m")
    r["LZ4/test1_legacy.lz4"] = (lambda: lz4_legacy(t1), "test1")
    r["LZ4/test5_legacy.lz4"] = (lambda: lz4_legacy(t5), "test5")
    r["LZ4/test6_legacy.lz4"] = (lambda: lz4_legacy(t6), "test6")
    r["LZ4/zeros.lz4"] = (lambda: lz4_legacy(bytes(18874368)), "zeros18m")
    return r


def main():
    manifest = {"answers": {"test1": {"literal": "Hello, World!\\n"}, "test5": {"zeros": 0}, "test6": {"zeros": 1 << 20},
                            "zeros5m": {"zeros": 5 * 1024 * 1024}, "zeros18m": {"zeros": 18874368}},
                "fixtures": {}}
    bad = 0
    for rel, (fn, answer) in recipes().items():
        try:
            data = fn()
        except Exception as e:  # noqa
            print(f"SKIP {rel}: {e}")
            continue
        sha = hashlib.sha256(data).hexdigest()
        oid, size = pointer(rel)
        ok = (sha == oid and len(data) == size)
        print(("OK   " if ok else "MISS ") + f"{rel} {len(data)} B sha256={sha[:16]}… pointer={oid[:16]}… size={size}")
        if not ok:
            bad += 1
            continue
        path = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(data)
        manifest["fixtures"][rel] = {"sha256": sha, "size": len(data), "answer": answer}
    # the one real (non-pointer) fixture in the reference tree: a 1 KiB LZ4 dictionary
    for a, content in (("test1", ANSWERS["test1"]), ("test6", ANSWERS["test6"])):
        oid, size = pointer(f"Answers/{a}.answer")
        assert hashlib.sha256(content).hexdigest() == oid and len(content) == size, a
    with open(os.path.join(OUT, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    print(f"{len(manifest['fixtures'])} fixtures written, {bad} not reproducible")
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""GzipHeader / ZlibHeader framing (swc_gzip_header_parse, swc_zlib_header_parse): hand-built headers with the result the
reference produces, read off the Swift source (Sources/GZip/GzipHeader.swift:68-199, Sources/Zlib/ZlibHeader.swift:47-93) —
NOT generated with the oracle, so product and oracle framing are pinned independently.  Pure host code: runs without a GPU."""
import struct
import zlib

import pytest

OK = "ok"
MAGIC, METHOD, FLAGS, HCRC = "GzipError.wrongMagic", "GzipError.wrongCompressionMethod", "GzipError.wrongFlags", "GzipError.wrongHeaderCRC"
TRAP = "EngineError.referenceTrap"


def fixed(flags, method=8, magic=b"\x1f\x8b", mtime=0, xfl=0, os_=3):
    return magic + bytes([method, flags]) + struct.pack("<I", mtime) + bytes([xfl, os_])


def with_hcrc(header):
    return header + struct.pack("<H", zlib.crc32(header) & 0xFFFF)


GZIP_CASES = [
    # (name, bytes, expected status, expected fields)
    ("minimal", fixed(0), OK, dict(header_len=10, name=None, comment=None, extra=[], text=False, mtime=0)),
    ("mtime+os", fixed(0, mtime=1482698300, os_=11), OK, dict(header_len=10, mtime=1482698300, os=11)),
    ("nine bytes", fixed(0)[:9], MAGIC, None),                                   # :70 bytesLeft >= 10
    ("bad magic", fixed(0, magic=b"\x1f\x8c"), MAGIC, None),                     # :75
    ("method 7", fixed(0, method=7), METHOD, None),                              # :81
    ("reserved flag 0x20", fixed(0x20), FLAGS, None),                            # :87
    ("reserved flag 0x80", fixed(0x80), FLAGS, None),
    ("ftext", fixed(0x01), OK, dict(text=True, header_len=10)),
    ("fextra: no xlen", fixed(0x04) + b"\x04", MAGIC, None),                     # :112 bytesLeft >= 2
    ("fextra: xlen 3", fixed(0x04) + b"\x03\x00" + b"ABC", MAGIC, None),         # :123 xlen >= 4
    ("fextra: xlen beyond data", fixed(0x04) + b"\x04\x00" + b"AB\x00", MAGIC, None),     # :123 bytesLeft >= xlen
    ("fextra: si2 zero", fixed(0x04) + b"\x04\x00" + b"A\x00\x00\x00", FLAGS, None),      # :131
    ("fextra: len > rest", fixed(0x04) + b"\x04\x00" + b"AB\x01\x00", MAGIC, None),       # :145 xlen(0) >= len(1)
    ("fextra: one field", fixed(0x04) + b"\x06\x00" + b"AB\x02\x00xy", OK, dict(extra=[(65, 66, b"xy")], header_len=18)),
    ("fextra: two fields", fixed(0x04) + b"\x0a\x00" + b"AB\x01\x00x" + b"CD\x01\x00y", OK,
     dict(extra=[(65, 66, b"x"), (67, 68, b"y")], header_len=22)),
    ("fextra: empty field", fixed(0x04) + b"\x04\x00" + b"AB\x00\x00", OK, dict(extra=[(65, 66, b"")], header_len=16)),
    # xlen 5: after the 4-byte field one byte of the area is left, the loop reads another field header across the area's end,
    # `xlen -= 4` goes negative and `guard xlen >= len` fails (:141-146)
    ("fextra: xlen 5 runs over", fixed(0x04) + b"\x05\x00" + b"AB\x00\x00" + b"XY\x00\x00\x00\x00", MAGIC, None),
    ("fextra: xlen 5, reserved id past the area", fixed(0x04) + b"\x05\x00" + b"AB\x00\x00" + b"X\x00\x00\x00\x00", FLAGS, None),
    ("fextra: xlen 5, data ends inside the second field header", fixed(0x04) + b"\x05\x00" + b"AB\x00\x00" + b"X", TRAP, None),
    ("fname", fixed(0x08) + b"abc\x00", OK, dict(name="abc", header_len=14)),
    ("fname latin-1", fixed(0x08) + b"caf\xe9\x00", OK, dict(name="café", header_len=15)),
    ("fname empty", fixed(0x08) + b"\x00", OK, dict(name="", header_len=11)),
    ("fname unterminated", fixed(0x08) + b"abc", MAGIC, None),                   # :162 isFinished
    ("fcomment", fixed(0x10) + b"hello\x00", OK, dict(comment="hello", name=None, header_len=16)),
    ("fcomment unterminated", fixed(0x10) + b"hello", MAGIC, None),              # :180
    ("fname+fcomment", fixed(0x18) + b"n\x00c\x00", OK, dict(name="n", comment="c", header_len=14)),
    ("fhcrc ok", with_hcrc(fixed(0x02)), OK, dict(header_len=12)),
    ("fhcrc wrong", fixed(0x02) + b"\x00\x00" if zlib.crc32(fixed(0x02)) & 0xFFFF else fixed(0x02) + b"\x01\x00", HCRC, None),
    ("fhcrc missing", fixed(0x02) + b"\x12", MAGIC, None),                       # :194 bytesLeft >= 2
    ("everything", with_hcrc(fixed(0x1F, mtime=7) + b"\x05\x00" + b"AB\x01\x00z" + b"name\x00" + b"cmt\x00"), OK,
     dict(text=True, mtime=7, extra=[(65, 66, b"z")], name="name", comment="cmt", header_len=10 + 2 + 5 + 5 + 4 + 2)),
    ("everything, crc over a damaged name", None, HCRC, None),
]


@pytest.fixture(scope="module")
def S():
    import swcompression_b200 as S
    return S


@pytest.mark.parametrize("name,data,status,fields", GZIP_CASES, ids=[c[0] for c in GZIP_CASES])
def test_gzip_header(S, name, data, status, fields):
    if data is None:
        good = with_hcrc(fixed(0x1F, mtime=7) + b"\x05\x00" + b"AB\x01\x00z" + b"name\x00" + b"cmt\x00")
        data = good.replace(b"name", b"nbme")
    for tail in (b"", b"\x03\x00" + b"\x00" * 8):                    # a Deflate payload behind the header changes nothing ...
        if tail and status in (MAGIC, TRAP) and name not in ("bad magic", "fextra: xlen 3", "fextra: len > rest", "fextra: xlen 5 runs over"):
            continue                                                  # ... except where the error IS running out of bytes
        if status == OK:
            h = S.GzipHeader(data + tail)
            want = dict(header_len=None, name=None, comment=None, extra=[], text=False, mtime=0, os=3)
            want.update(fields)
            assert h.fileName == want["name"] and h.comment == want["comment"] and h.isTextFile == want["text"]
            assert [(f.si1, f.si2, f.bytes) for f in h.extraFields] == want["extra"]
            assert (0 if h.modificationTime is None else int(h.modificationTime.timestamp())) == want["mtime"]
            assert h.osType == {3: "unix", 11: "ntfs"}[want["os"]] and h.compressionMethod == "deflate"
            if want["header_len"] is not None:
                assert h.headerLength == want["header_len"]
        else:
            with pytest.raises(S.SWCompressionError) as e:
                S.GzipHeader(data + tail)
            assert str(e.value) == status, (name, str(e.value))


def test_gzip_header_at_member_offset(S):
    a = fixed(0x08) + b"first\x00" + b"\x03\x00" + b"\x00" * 8
    b = fixed(0x08, mtime=9) + b"second\x00" + b"\x03\x00" + b"\x00" * 8
    h = S.GzipHeader(a + b, _member_off=len(a))
    assert h.fileName == "second" and int(h.modificationTime.timestamp()) == 9


ZLIB_CASES = [
    ("one byte", b"\x78", "ZlibError.wrongCompressionMethod", None),                         # :49
    ("default", b"\x78\x9c", OK, ("defaultAlgorithm", 32768)),
    ("fastest, 256-byte window", bytes([0x08, 0x1d]), OK, ("fastestAlgorithm", 256)),        # 0x081d % 31 == 0
    ("slow", b"\x78\xda", OK, ("slowAlgorithm", 32768)),
    ("method 7", b"\x77\x9c", "ZlibError.wrongCompressionMethod", None),                     # :57
    ("cinfo 8", b"\x88\x9c", "ZlibError.wrongCompressionInfo", None),                        # :63 (before the FCHECK test)
    ("fcheck", b"\x78\x9d", "ZlibError.wrongFcheck", None),                                  # :84
    ("fdict without the dictionary id", b"\x78\x20\x00\x00\x00", "ZlibError.wrongFcheck", None),   # :89 bytesLeft >= 4
    ("fdict", b"\x78\x20\x00\x00\x00\x00", OK, ("fastestAlgorithm", 32768)),
]


@pytest.mark.parametrize("name,data,status,fields", ZLIB_CASES, ids=[c[0] for c in ZLIB_CASES])
def test_zlib_header(S, name, data, status, fields):
    assert (0x7820 % 31 == 0) and (0x081d % 31 == 0)
    if status == OK:
        z = S.ZlibHeader(data)
        assert (z.compressionLevel, z.windowSize) == fields and z.compressionMethod == "deflate"
    else:
        with pytest.raises(S.SWCompressionError) as e:
            S.ZlibHeader(data)
        assert str(e.value) == status

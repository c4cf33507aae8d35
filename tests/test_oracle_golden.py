"""Pins the CPU oracle (oracle/) against the reference's own golden vectors:
fixture -> answer tests, inline malformed-bitstream vectors, known-answer hashes, and differential runs against
zlib / bz2 / lzma / liblz4 on valid streams.  CPU only."""
import bz2
import gzip
import hashlib
import lzma
import random
import struct
import zlib

import pytest

import helpers as H
import oracle_xxh

S_OK = 0


# ---------------------------------------------------------------- fixture -> answer (Tests/*Tests.swift)
@pytest.mark.parametrize("rel,ans", H.fixtures("Deflate/"))
def test_deflate_fixtures(oracle, rel, ans):
    st, out, used = oracle.deflate_decompress(H.fixture(rel))
    assert st == S_OK and out == H.answer(ans)


@pytest.mark.parametrize("rel,ans", H.fixtures("GZip/"))
def test_gzip_fixtures(oracle, rel, ans):                       # GzipTests.swift:76-138
    st, out, _ = oracle.gzip_unarchive(H.fixture(rel))
    assert st == S_OK and out == H.answer(ans)


def test_zlib_fixtures(oracle):                                 # ZlibTests.swift:13-30
    st, out, _ = oracle.zlib_unarchive(H.fixture("Zlib/test_empty.zlib"))
    assert st == S_OK and out == b""
    st, out, _ = oracle.zlib_unarchive(H.fixture("Zlib/test.zlib"))   # header only: Deflate sees < 10 bits
    assert st == 102


@pytest.mark.parametrize("rel,ans", H.fixtures("BZip2/"))
def test_bzip2_fixtures(oracle, rel, ans):                      # BZip2Tests.swift:21-59
    st, out, _ = oracle.bzip2_decompress(H.fixture(rel))
    assert st == S_OK and out == H.answer(ans)


@pytest.mark.parametrize("rel,ans", H.fixtures("XZ/"))
def test_xz_fixtures(oracle, rel, ans):                         # XzTests.swift:21-56
    st, out, _ = oracle.xz_unarchive(H.fixture(rel))
    assert st == S_OK and out == H.answer(ans)


def test_lzma_fixture(oracle):                                  # LzmaTests.swift:21-40
    st, out, _ = oracle.lzma_decompress(H.fixture("LZMA/test_empty.lzma"))
    assert st == S_OK and out == b""


@pytest.mark.parametrize("rel,ans", H.fixtures("LZ4/"))
def test_lz4_fixtures(oracle, rel, ans):                        # LZ4Tests.swift:32-138 (frames, legacy, B4-B7, *_BD)
    st, out, _ = oracle.lz4_decompress(H.fixture(rel))
    assert st == S_OK and out == H.answer(ans)


# ---------------------------------------------------------------- inline vectors (DeflateTests.swift:35-194)
def _all_zero_lengths_writer(first):
    w = H.LsbBitWriter()
    w.write_bits([1, 0, 1]); w.write_number(29, 5); w.write_number(1, 5); w.write_number(14, 4)
    w.write_number(0, 3); w.write_number(3, 3); w.write_number(2, 3)
    for _ in range(10):
        w.write_number(0, 3)
    for v in (2, 0, 3, 0, 2):
        w.write_number(v, 3)
    if first:      # :127-146 all lit/len lengths zero, two distance codes
        for v, c in ((1, 2), (127, 7), (1, 2), (127, 7), (7, 3), (7, 3), (0, 2), (0, 2)):
            w.write_number(v, c)
    else:          # :170-190 empty distance tree, literal 0 + EOB
        for v, c in ((3, 3), (1, 2), (127, 7), (1, 2), (106, 7), (2, 2), (1, 2), (20, 7), (0, 2), (2, 3)):
            w.write_number(v, c)
    return w.data


DEFLATE_INLINE = [
    (bytes([0b0000_0101, 0, 0b1010_0010, 0b0000_1101]), None),                                    # :46 symbol 16 first
    (bytes([0b0000_0101, 0, 0b1010_0010, 0b1110_1101, 0xFF, 0xFF, 0b0000_0001]), None),           # :65
    (bytes([0b0000_0101, 0, 0b1010_0010, 0b1110_1101, 0xFF, 0b1011_0011, 0b0000_0101]), None),    # :80
    (bytes([0b0000_0101, 0, 0, 0]), None),                                                        # :97
    (_all_zero_lengths_writer(True), None),                                                       # :146
    (_all_zero_lengths_writer(False), b"\x00"),                                                   # :193
]


@pytest.mark.parametrize("data,expect", DEFLATE_INLINE)
def test_deflate_inline_vectors(oracle, data, expect):
    st, out, _ = oracle.deflate_decompress(data)
    if expect is None:
        assert st != S_OK
    else:
        assert st == S_OK and out == expect


def test_deflate_inline_bytes_match_survey():
    assert _all_zero_lengths_writer(True).hex() == "edc13101000000c2a0fefd7f00"
    assert _all_zero_lengths_writer(False).hex() == "edc13101000000c2a0f54f6d1404"


def test_zlib_rejects_what_reference_accepts():
    # zlib is not a sufficient oracle: it rejects the all-zero-distance-tree vector the reference decodes to [0].
    d = zlib.decompressobj(-15)
    with pytest.raises(zlib.error):
        d.decompress(_all_zero_lengths_writer(False))


def test_short_inputs_throw(oracle):
    # LzmaTests.swift:44-46, BZip2Tests.swift:62,78, LZ4Tests.swift:87-108, XzTests.swift:112-115, ZlibTests.swift:45-48, GzipTests.swift:174-178
    for n in range(0, 20):
        junk = bytes(range(n))
        assert oracle.lzma_decompress(junk)[0] != S_OK
        assert oracle.bzip2_decompress(junk)[0] != S_OK
        assert oracle.xz_unarchive(junk)[0] != S_OK or n == 0
        assert oracle.gzip_unarchive(junk)[0] != S_OK
        assert oracle.zlib_unarchive(junk)[0] != S_OK
    assert oracle.lz4_decompress(b"")[0] == 501 and oracle.lz4_decompress(b"\0")[0] == 501
    assert oracle.lz4_decompress(bytes(1 << 20))[0] == 502


# ---------------------------------------------------------------- known-answer hashes
XXH = [(b"", 0x02cc5d05), (b"a", 0x550d7456), (b"abc", 0x32d153ff), (b"message digest", 0x7c948494),
       (b"abcdefghijklmnopqrstuvwxyz", 0x63a14d5f),
       (b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789", 0x9c285e64),
       (b"1234567890" * 8, 0x9c05f475)]


@pytest.mark.parametrize("msg,h", XXH)
def test_xxh32_kat(oracle, msg, h):                             # XxHash32Tests.swift:12-59
    assert oracle.xxh32(msg) == h
    assert oracle_xxh.xxh32(msg) == h


@pytest.mark.parametrize("msg", [b"", b"a", b"abc", b"message digest", b"abcdefghijklmnopqrstuvwxyz",
                                 b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789", b"1234567890" * 8,
                                 bytes(55), bytes(56), bytes(63), bytes(64), bytes(119), bytes(1000)])
def test_sha256_kat(oracle, msg):                               # Sha256Tests.swift:12-66 (FIPS 180-4 answers)
    assert oracle.sha256(msg) == hashlib.sha256(msg).digest()


def test_crcs(oracle):
    rng = random.Random(7)
    for n in (0, 1, 3, 255, 4096, 70001):
        b = bytes(rng.getrandbits(8) for _ in range(n))
        assert oracle.crc32(b) == zlib.crc32(b)
        assert oracle.adler32(b) == zlib.adler32(b)
    assert oracle.crc64(b"123456789") == 0x995DC9BBDF1939FA       # CRC-64/XZ check value
    assert oracle.bzip2_crc32(b"123456789") == 0xFC891918          # CRC-32/BZIP2 check value


# ---------------------------------------------------------------- round trips through conformant compressors
@pytest.mark.parametrize("raw", H.ROUNDTRIP_STRINGS)
def test_roundtrip_strings(oracle, raw):
    for lvl in (0, 1, 6, 9):
        assert oracle.deflate_decompress(H.raw_deflate(raw, lvl, 8))[:2] == (S_OK, raw)
    assert oracle.zlib_unarchive(zlib.compress(raw))[:2] == (S_OK, raw)
    assert oracle.gzip_unarchive(gzip.compress(raw))[:2] == (S_OK, raw)
    assert oracle.bzip2_decompress(bz2.compress(raw))[:2] == (S_OK, raw)
    assert oracle.xz_unarchive(lzma.compress(raw))[:2] == (S_OK, raw)
    assert oracle.lzma_decompress(lzma.compress(raw, format=lzma.FORMAT_ALONE))[:2] == (S_OK, raw)
    if raw:
        assert oracle.lz4_block(H.lz4_block_compress(raw))[:2] == (S_OK, raw)
    assert oracle.lz4_decompress(H.lz4_frame_independent([raw] if raw else [], content_checksum=True, block_checksum=True))[:2] == (S_OK, raw)


@pytest.mark.parametrize("seed", range(4))
def test_differential_textlike(oracle, seed):
    raw = H.textlike(65536 + seed * 1000, seed)
    comp = H.raw_deflate(raw)
    st, out, used = oracle.deflate_decompress(comp)
    assert (st, out) == (S_OK, raw) and (used + 7) // 8 == len(comp)
    assert oracle.bzip2_decompress(bz2.compress(raw, 9))[:2] == (S_OK, raw)
    xz = lzma.compress(raw, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC64, filters=[{"id": lzma.FILTER_LZMA2, "preset": 6, "dict_size": 1 << 20}])
    assert oracle.xz_unarchive(xz)[:2] == (S_OK, raw)
    assert oracle.lz4_block(H.lz4_block_compress(raw))[:2] == (S_OK, raw)
    # delta + lzma2 filter chain (XzTests test_delta_filter analogue)
    xzd = lzma.compress(raw, format=lzma.FORMAT_XZ, filters=[{"id": lzma.FILTER_DELTA, "dist": 3}, {"id": lzma.FILTER_LZMA2, "preset": 1}])
    assert oracle.xz_unarchive(xzd)[:2] == (S_OK, raw)


def test_multi_member_and_stream(oracle):
    a, b = H.textlike(5000, 11), H.textlike(7000, 12)
    st, parts, _ = oracle.gzip_multi_unarchive(gzip.compress(a) + gzip.compress(b))
    assert st == S_OK and parts == [a, b]
    st, parts, _ = oracle.bzip2_multi_decompress(bz2.compress(a) + bz2.compress(b))
    assert st == S_OK and parts == [a, b]
    st, parts, whole = oracle.xz_split_unarchive(lzma.compress(a) + bytes(4) + lzma.compress(b))
    assert st == S_OK and parts == [a, b]
    f1 = H.lz4_frame_independent([a]); f2 = H.lz4_frame_independent([b])
    skip = struct.pack("<II", 0x184D2A53, 5) + b"hello"
    st, parts, _ = oracle.lz4_multi_decompress(f1 + skip + f2)
    assert st == S_OK and parts == [a, b]


# ---------------------------------------------------------------- truncation fuzz: must fail, must not crash
def test_truncation_fuzz(oracle):
    rng = random.Random(99)
    raw = H.textlike(20000, 5)
    cases = [
        (oracle.deflate_decompress, H.raw_deflate(raw)),
        (oracle.deflate_decompress, H.raw_deflate(raw, 0)),          # stored blocks
        (oracle.deflate_decompress, zlib.compress(raw, 6, wbits=-15) if hasattr(zlib, "Z_FIXED") else H.raw_deflate(raw)),
        (oracle.gzip_unarchive, gzip.compress(raw)),
        (oracle.zlib_unarchive, zlib.compress(raw)),
        (oracle.bzip2_decompress, bz2.compress(raw)),
        (oracle.lz4_decompress, H.lz4_frame_independent([raw], content_checksum=True)),
        (oracle.xz_unarchive, lzma.compress(raw)),
        (oracle.lzma_decompress, lzma.compress(raw, format=lzma.FORMAT_ALONE)),
    ]
    for fn, data in cases:
        for _ in range(25):
            cut = rng.randrange(1, len(data))
            st = fn(data[:cut])[0]
            assert st != S_OK, (fn.__name__, cut)


# ---------------------------------------------------------------- checksum mismatch still returns the payload
def test_checksum_mismatch_payload(oracle):
    raw = H.textlike(3000, 3)
    g = bytearray(gzip.compress(raw)); g[-8] ^= 1
    assert oracle.gzip_unarchive(bytes(g))[:2] == (605, raw)                     # GzipTests.swift:190-228
    z = bytearray(zlib.compress(raw)); z[-1] ^= 1
    assert oracle.zlib_unarchive(bytes(z))[:2] == (705, raw)                     # ZlibTests.swift:59-75
    b = bytearray(bz2.compress(raw)); b[10] ^= 1                                 # block CRC byte
    assert oracle.bzip2_decompress(bytes(b))[:2] == (210, raw)                   # BZip2Tests.swift:81-97
    x = bytearray(lzma.compress(raw, check=lzma.CHECK_CRC32))
    st, out, _ = oracle.xz_unarchive(bytes(x)); assert st == S_OK
    # flip a byte inside the block check (the 4 bytes before the index indicator)
    idx = len(x) - 12 - 8 - 4
    for off in range(idx - 8, idx + 1):
        y = bytearray(x); y[off] ^= 0xFF
        st, out, _ = oracle.xz_unarchive(bytes(y))
        if st == 807:
            assert out == raw
            break
    else:
        pytest.fail("no wrongCheck produced")
    l = bytearray(H.lz4_frame_independent([raw], content_checksum=True)); l[-1] ^= 1
    assert oracle.lz4_decompress(bytes(l))[:2] == (503, raw)                     # LZ4Tests.swift:186-203

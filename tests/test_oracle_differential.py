"""Oracle vs independent implementations (zlib / bz2 / liblzma / gzip) over the encoder option space.

The oracle restates the Swift reference; these runs pin it from the other side: every stream a conformant encoder can emit —
all zlib strategies, levels, memory levels and window sizes, flush-split multi-block streams, bzip2 block sizes and run-length
edge cases, LZMA lc/lp/pb grids, XZ check types and filter chains, gzip optional header fields — must decode to the original
bytes with the exact consumed size.  CPU only; the GPU suite compares the CUDA path with this oracle."""
import bz2
import gzip
import io
import lzma
import random
import struct
import zlib

import pytest

import helpers as H

S_OK = 0


def _corpora(rng):
    text = H.textlike(40000, 7)
    yield "text", text
    yield "zeros", bytes(50000)
    yield "random", bytes(rng.randrange(256) for _ in range(20000))
    yield "period3", (b"abc" * 20000)[:50001]
    yield "period258", bytes((i * 7) & 0xFF for i in range(258)) * 150
    yield "runs", b"".join(bytes([rng.randrange(4)]) * rng.choice((1, 2, 3, 4, 5, 254, 255, 256, 259, 300)) for _ in range(600))
    yield "short", b"a"
    yield "empty", b""


def test_deflate_strategy_level_memlevel_window_grid(oracle):
    rng = random.Random(101)
    strategies = (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED)
    n = 0
    for name, raw in _corpora(rng):
        for strat in strategies:
            for level, mem, wbits in ((0, 8, 15), (1, 1, 9), (1, 9, 15), (4, 5, 12), (6, 8, 15), (9, 9, 15), (9, 1, 10)):
                c = zlib.compressobj(level, zlib.DEFLATED, -wbits, mem, strat)
                comp = c.compress(raw) + c.flush()
                st, out, used = oracle.deflate_decompress(comp)
                assert (st, out) == (S_OK, raw), (name, strat, level, mem, wbits)
                assert (used + 7) // 8 == len(comp), (name, strat, level, mem, wbits, used, len(comp))
                n += 1
    assert n == 8 * 5 * 7


def test_deflate_flush_split_streams(oracle):
    """sync / full / block flushes put empty stored blocks and block boundaries at arbitrary bit offsets"""
    rng = random.Random(102)
    raw = H.textlike(120000, 8)
    for trial in range(12):
        c = zlib.compressobj(rng.choice((1, 6, 9)), zlib.DEFLATED, -15, rng.choice((1, 8, 9)))
        comp, pos = b"", 0
        while pos < len(raw):
            step = rng.randrange(1, 9000)
            comp += c.compress(raw[pos:pos + step])
            comp += c.flush(rng.choice((zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_BLOCK, zlib.Z_NO_FLUSH, zlib.Z_PARTIAL_FLUSH)))
            pos += step
        comp += c.flush()
        st, out, used = oracle.deflate_decompress(comp)
        assert (st, out) == (S_OK, raw) and (used + 7) // 8 == len(comp)


def test_gzip_optional_header_fields(oracle):
    """FEXTRA / FNAME / FCOMMENT / FHCRC in every combination (GzipHeader.swift:68-199); python's gzip reads the same bytes"""
    raw = H.textlike(3000, 9)
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = c.compress(raw) + c.flush()
    trailer = struct.pack("<II", zlib.crc32(raw), len(raw))
    for flags in range(0, 32, 2):                                    # bit 0 (FTEXT) carries no field
        for ftext in (0, 1):
            f = flags | ftext
            hdr = bytes([0x1F, 0x8B, 8, f]) + struct.pack("<I", 1234567) + bytes([2, 3])
            if f & 4:
                sub = b"AP" + struct.pack("<H", 5) + b"hello"
                hdr += struct.pack("<H", len(sub)) + sub
            if f & 8:
                hdr += b"name.txt\x00"
            if f & 16:
                hdr += b"a comment\x00"
            if f & 2:
                hdr += struct.pack("<H", zlib.crc32(hdr) & 0xFFFF)
            blob = hdr + body + trailer
            assert gzip.GzipFile(fileobj=io.BytesIO(blob)).read() == raw
            st, out, _ = oracle.gzip_unarchive(blob)
            assert (st, out) == (S_OK, raw), f
            st, parts, _ = oracle.gzip_multi_unarchive(blob + blob)
            assert st == S_OK and parts == [raw, raw]


def test_zlib_levels_and_dictionary_flag(oracle):
    raw = H.textlike(20000, 10)
    for level in range(10):
        for wbits in (9, 12, 15):
            c = zlib.compressobj(level, zlib.DEFLATED, wbits)
            comp = c.compress(raw) + c.flush()
            assert oracle.zlib_unarchive(comp)[:2] == (S_OK, raw)


def test_bzip2_block_sizes_and_run_edges(oracle):
    rng = random.Random(103)
    for name, raw in _corpora(rng):
        for level in (1, 5, 9):
            comp = bz2.compress(raw, level)
            st, out, used = oracle.bzip2_decompress(comp)
            assert (st, out) == (S_OK, raw), (name, level)
            assert (used + 7) // 8 == len(comp)
    # RLE1 boundaries: runs of exactly 4..7, 255+4, and a run cut by the 100 KB block boundary
    for run in (3, 4, 5, 6, 7, 258, 259, 260, 1000):
        raw = b"x" + b"y" * run + b"z" + b"\x00" * run + H.textlike(500, run)
        assert oracle.bzip2_decompress(bz2.compress(raw, 1))[:2] == (S_OK, raw)
    raw = H.textlike(99990, 11) + b"q" * 40 + H.textlike(50000, 12)
    assert oracle.bzip2_decompress(bz2.compress(raw, 1))[:2] == (S_OK, raw)
    # all 256 byte values in use / only one value in use
    raw = bytes(range(256)) * 40
    assert oracle.bzip2_decompress(bz2.compress(raw))[:2] == (S_OK, raw)
    assert oracle.bzip2_decompress(bz2.compress(b"\xff" * 100000, 1))[:2] == (S_OK, b"\xff" * 100000)


def test_lzma_alone_lc_lp_pb_grid(oracle):
    raw = H.textlike(30000, 13) + bytes(3000) + bytes(random.Random(104).randrange(256) for _ in range(3000))
    for lc in range(5):
        for lp in range(5 - lc):
            for pb in (0, 2, 4):
                f = [{"id": lzma.FILTER_LZMA1, "lc": lc, "lp": lp, "pb": pb, "dict_size": 1 << 16, "mode": lzma.MODE_NORMAL, "nice_len": 64, "mf": lzma.MF_BT4}]
                comp = lzma.compress(raw, format=lzma.FORMAT_ALONE, filters=f)
                st, out, _ = oracle.lzma_decompress(comp)
                if pb == 4 and st == 2:
                    # faithful to a reference defect: `probabilities` has 2*192 + 48 = 432 entries but isRep0Long is indexed
                    # 241 + (state << 4) + posState, which reaches 432 for state 11 / posState 15 (LZMADecoder.swift:87,187):
                    # the Swift array access traps on such (valid) streams -> SWC_ERR_REFERENCE_TRAP in oracle and product
                    continue
                assert (st, out) == (S_OK, raw), (lc, lp, pb)


def test_xz_checks_presets_and_filter_chains(oracle):
    raw = H.textlike(60000, 14) + bytes(5000)
    for check in (lzma.CHECK_NONE, lzma.CHECK_CRC32, lzma.CHECK_CRC64, lzma.CHECK_SHA256):
        for preset in (0, 3, 6, 9 | lzma.PRESET_EXTREME):
            comp = lzma.compress(raw, format=lzma.FORMAT_XZ, check=check, preset=preset)
            assert oracle.xz_unarchive(comp)[:2] == (S_OK, raw), (check, preset)
    for dist in (1, 2, 4, 256):
        chain = [{"id": lzma.FILTER_DELTA, "dist": dist}, {"id": lzma.FILTER_LZMA2, "preset": 2, "dict_size": 1 << 16}]
        comp = lzma.compress(raw, format=lzma.FORMAT_XZ, filters=chain)
        assert oracle.xz_unarchive(comp)[:2] == (S_OK, raw), dist
    # incompressible input: LZMA2 falls back to uncompressed chunks (control bytes 1 / 2)
    rnd = bytes(random.Random(105).randrange(256) for _ in range(200000))
    assert oracle.xz_unarchive(lzma.compress(rnd, preset=1))[:2] == (S_OK, rnd)
    # several streams with stream padding between them
    a, b, c = raw[:1000], raw[1000:30000], b""
    blob = lzma.compress(a) + bytes(8) + lzma.compress(b, check=lzma.CHECK_SHA256) + lzma.compress(c)
    st, parts, _ = oracle.xz_split_unarchive(blob)
    assert st == S_OK and parts == [a, b, c]
    # multi-block stream: two independently encoded streams cannot be merged by liblzma's python binding, but a block-split
    # encoder (xz -T) writes the same container; emulate it with LZMA2 raw blocks is out of reach here, so the multi-block
    # path stays pinned by the reference's own fixture (test4.xz family) in test_oracle_golden.py.


def test_lzma2_raw_dictionary_sizes(oracle):
    raw = H.textlike(150000, 15)
    for bits in (12, 16, 20, 24):
        f = [{"id": lzma.FILTER_LZMA2, "preset": 4, "dict_size": 1 << bits}]
        comp = lzma.compress(raw, format=lzma.FORMAT_RAW, filters=f)
        # LZMA2Decoder.swift:17-30: the dictionary-size byte precedes the chunks in SWCompression's LZMA2.decompress(data:)
        for dict_byte in range(41):
            size = (2 | (dict_byte & 1)) << (dict_byte // 2 + 11) if dict_byte < 40 else 0xFFFFFFFF
            if size >= (1 << bits):
                break
        st, out, _ = oracle.lzma2_decompress(bytes([dict_byte]) + comp)
        assert (st, out) == (S_OK, raw), bits

"""CUDA path over the encoder option space of tests/test_oracle_differential.py: every stream a conformant encoder emits
(zlib strategies x levels x memLevels x windows, flush-split streams, gzip header fields, bzip2 block sizes / run edges, LZMA
lc/lp/pb grid, XZ checks / presets / delta chains) through the C ABI, compared with the original bytes and with the oracle's
consumed size / status.  Deflate runs on both Huffman kernels: the warp-per-unit one (small batch) and the thread-per-unit one
(the same streams tiled to >= 20000 units)."""
import bz2
import lzma
import random
import zlib

import numpy as np
import pytest

import helpers as H
from test_oracle_differential import _corpora

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available()
    import swcompression_b200 as S
    return S


def _run_batch(units, cap):
    from swcompression_b200.batch import Batch
    b = Batch.from_units("deflate", units, cap)
    b.run()
    st, ln, used = b.results()
    return st, ln, used, b


def test_deflate_option_grid_on_both_kernels(oracle):
    rng = random.Random(101)
    strategies = (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED)
    units, raws = [], []
    for name, raw in _corpora(rng):
        for strat in strategies:
            for level, mem, wbits in ((0, 8, 15), (1, 1, 9), (1, 9, 15), (4, 5, 12), (6, 8, 15), (9, 9, 15), (9, 1, 10)):
                c = zlib.compressobj(level, zlib.DEFLATED, -wbits, mem, strat)
                units.append(c.compress(raw) + c.flush()); raws.append(raw)
    raw = H.textlike(120000, 8)                                   # flush-split streams: empty stored blocks, odd bit offsets
    for trial in range(12):
        c = zlib.compressobj(rng.choice((1, 6, 9)), zlib.DEFLATED, -15, rng.choice((1, 8, 9)))
        comp, pos = b"", 0
        while pos < len(raw):
            step = rng.randrange(1, 9000)
            comp += c.compress(raw[pos:pos + step])
            comp += c.flush(rng.choice((zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH, zlib.Z_BLOCK, zlib.Z_NO_FLUSH, zlib.Z_PARTIAL_FLUSH)))
            pos += step
        units.append(comp + c.flush()); raws.append(raw)
    cap = 120000
    # (a) warp-per-unit kernel: one batch of the distinct streams
    st, ln, used, b = _run_batch(units, cap)
    outs = b.outputs()
    for i, u in enumerate(units):
        assert st[i] == 0 and outs[i] == raws[i], i
        assert (int(used[i]) + 7) // 8 == len(u), i
    # (b) thread-per-unit kernel: >= 20000 units (only the streams that fit a 64 KiB fence, to keep the batch near 1 GB)
    small = [i for i, r in enumerate(raws) if len(r) <= 50001]
    order = [small[k % len(small)] for k in range(20480)]
    st, ln, used, b = _run_batch([units[i] for i in order], 50016)
    assert (st == 0).all()
    want_len = np.array([len(raws[i]) for i in order])
    assert (ln == want_len).all()
    want_used = np.array([len(units[i]) for i in order])
    assert (((used + 7) // 8) == want_used).all()
    outs = b.outputs()
    for j in range(0, len(order), 97):
        assert outs[j] == raws[order[j]], (j, order[j])
    for j in range(len(small)):                                   # every distinct stream at least once
        assert outs[j] == raws[order[j]], (j, order[j])


def test_bzip2_block_sizes_and_run_edges(gpu, oracle):
    rng = random.Random(103)
    for name, raw in _corpora(rng):
        for level in (1, 5, 9):
            assert gpu.BZip2.decompress(bz2.compress(raw, level)) == raw, (name, level)
    for run in (3, 4, 5, 6, 7, 258, 259, 260, 1000):
        raw = b"x" + b"y" * run + b"z" + b"\x00" * run + H.textlike(500, run)
        assert gpu.BZip2.decompress(bz2.compress(raw, 1)) == raw, run
    raw = H.textlike(99990, 11) + b"q" * 40 + H.textlike(50000, 12)
    assert gpu.BZip2.decompress(bz2.compress(raw, 1)) == raw
    raw = bytes(range(256)) * 40
    assert gpu.BZip2.decompress(bz2.compress(raw)) == raw
    assert gpu.BZip2.decompress(bz2.compress(b"\xff" * 100000, 1)) == b"\xff" * 100000


def test_lzma_grid_and_xz_options(gpu, oracle):
    raw = H.textlike(30000, 13) + bytes(3000) + bytes(random.Random(104).randrange(256) for _ in range(3000))
    for lc in range(5):
        for lp in range(5 - lc):
            for pb in (0, 2, 4):
                f = [{"id": lzma.FILTER_LZMA1, "lc": lc, "lp": lp, "pb": pb, "dict_size": 1 << 16, "mode": lzma.MODE_NORMAL, "nice_len": 64, "mf": lzma.MF_BT4}]
                comp = lzma.compress(raw, format=lzma.FORMAT_ALONE, filters=f)
                ost, oout, _ = oracle.lzma_decompress(comp)
                if ost == 0:
                    assert gpu.LZMA.decompress(comp) == raw, (lc, lp, pb)
                else:                                            # the reference's 432-entry probabilities defect (pb = 4)
                    with pytest.raises(gpu.SWCompressionError) as e:
                        gpu.LZMA.decompress(comp)
                    assert e.value.code == ost, (lc, lp, pb)
    raw = H.textlike(60000, 14) + bytes(5000)
    for check in (lzma.CHECK_NONE, lzma.CHECK_CRC32, lzma.CHECK_CRC64, lzma.CHECK_SHA256):
        for preset in (0, 3, 6, 9 | lzma.PRESET_EXTREME):
            assert gpu.XZArchive.unarchive(lzma.compress(raw, format=lzma.FORMAT_XZ, check=check, preset=preset)) == raw, (check, preset)
    for dist in (1, 2, 4, 256):
        chain = [{"id": lzma.FILTER_DELTA, "dist": dist}, {"id": lzma.FILTER_LZMA2, "preset": 2, "dict_size": 1 << 16}]
        assert gpu.XZArchive.unarchive(lzma.compress(raw, format=lzma.FORMAT_XZ, filters=chain)) == raw, dist
    rnd = bytes(random.Random(105).randrange(256) for _ in range(200000))
    assert gpu.XZArchive.unarchive(lzma.compress(rnd, preset=1)) == rnd
    a, b, c = raw[:1000], raw[1000:30000], b""
    blob = lzma.compress(a) + bytes(8) + lzma.compress(b, check=lzma.CHECK_SHA256) + lzma.compress(c)
    assert gpu.XZArchive.splitUnarchive(blob) == [a, b, c]

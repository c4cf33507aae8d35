"""Parity of the CUDA LZMA / LZMA2 / XZ path (through the C ABI) with the CPU oracle."""
import lzma
import random

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available()
    import swcompression_b200 as S
    return S


@pytest.mark.parametrize("rel,ans", H.fixtures("XZ/"))
def test_xz_fixtures(gpu, rel, ans):
    assert gpu.XZArchive.unarchive(H.fixture(rel)) == H.answer(ans)


def test_lzma_fixture(gpu):
    assert gpu.LZMA.decompress(H.fixture("LZMA/test_empty.lzma")) == b""


@pytest.mark.parametrize("raw", H.ROUNDTRIP_STRINGS)
def test_roundtrip_strings(gpu, raw):
    assert gpu.XZArchive.unarchive(lzma.compress(raw)) == raw
    assert gpu.LZMA.decompress(lzma.compress(raw, format=lzma.FORMAT_ALONE)) == raw


def test_short_inputs(gpu, oracle):
    for n in range(0, 20):
        junk = bytes(range(n))
        for fn, ofn in ((gpu.LZMA.decompress, oracle.lzma_decompress), (gpu.XZArchive.unarchive, oracle.xz_unarchive),
                        (gpu.LZMA2.decompress, oracle.lzma2_decompress)):
            ost, oout, _ = ofn(junk)
            try:
                out = fn(junk)
                assert ost == 0 and out == oout
            except gpu.SWCompressionError as e:
                assert e.code == ost, (fn, n, e.code, ost)


def test_config5_shape_1mib_dict_stream(gpu, oracle):
    raw = H.textlike(1 << 20, 5)
    xz = lzma.compress(raw, format=lzma.FORMAT_XZ, check=lzma.CHECK_CRC64, filters=[{"id": lzma.FILTER_LZMA2, "preset": 6, "dict_size": 1 << 20}])
    assert oracle.xz_unarchive(xz)[:2] == (0, raw)
    assert gpu.XZArchive.unarchive(xz) == raw


def test_variants(gpu, oracle):
    raw = H.textlike(200000, 51)
    for check in (lzma.CHECK_NONE, lzma.CHECK_CRC32, lzma.CHECK_CRC64, lzma.CHECK_SHA256):
        assert gpu.XZArchive.unarchive(lzma.compress(raw, check=check)) == raw
    xzd = lzma.compress(raw, format=lzma.FORMAT_XZ, filters=[{"id": lzma.FILTER_DELTA, "dist": 3}, {"id": lzma.FILTER_LZMA2, "preset": 1}])
    assert gpu.XZArchive.unarchive(xzd) == raw
    for lc, lp, pb in ((0, 0, 0), (3, 0, 2), (1, 3, 4), (4, 0, 0), (0, 4, 1), (2, 2, 3)):   # liblzma only encodes lc+lp <= 4
        comp = lzma.compress(raw, format=lzma.FORMAT_ALONE, filters=[{"id": lzma.FILTER_LZMA1, "lc": lc, "lp": lp, "pb": pb, "dict_size": 1 << 16}])
        ost, oout, _ = oracle.lzma_decompress(comp)
        assert ost == 0 and oout == raw
        assert gpu.LZMA.decompress(comp) == raw
    # raw LZMA1 stream with explicit properties and known size (LZMA.decompress(data:properties:uncompressedSize:))
    rawstream = lzma.compress(raw, format=lzma.FORMAT_RAW, filters=[{"id": lzma.FILTER_LZMA1, "lc": 3, "lp": 0, "pb": 2, "dict_size": 1 << 20}])
    props = gpu.LZMAProperties(lc=3, lp=0, pb=2, dictionarySize=1 << 20)
    assert gpu.LZMA.decompress(rawstream, props, len(raw)) == raw
    # raw LZMA2 with the dictionary byte in front (LZMA2.decompress(data:))
    l2 = lzma.compress(raw, format=lzma.FORMAT_RAW, filters=[{"id": lzma.FILTER_LZMA2, "preset": 6, "dict_size": 1 << 20}])
    assert gpu.LZMA2.decompress(bytes([18]) + l2) == raw
    a, b = H.textlike(5000, 52), bytes(70000)
    assert gpu.XZArchive.splitUnarchive(lzma.compress(a) + bytes(4) + lzma.compress(b)) == [a, b]
    assert gpu.XZArchive.unarchive(lzma.compress(a) + lzma.compress(b)) == a + b


def test_incompressible_uses_stored_chunks(gpu):
    rng = random.Random(3)
    raw = bytes(rng.getrandbits(8) for _ in range(150000))
    assert gpu.XZArchive.unarchive(lzma.compress(raw)) == raw


def test_lzma2_batch(oracle):
    from swcompression_b200.batch import Batch
    raws = [H.textlike(50000 + 1000 * i, 600 + i) for i in range(24)]
    units = [lzma.compress(r, format=lzma.FORMAT_RAW, filters=[{"id": lzma.FILTER_LZMA2, "preset": 6, "dict_size": 1 << 20}]) for r in raws]
    b = Batch.from_units("lzma2", units, 131072, aux=bytes([18] * len(units)))
    b.run()
    st, ln, used = b.results()
    outs = b.outputs()
    for i, u in enumerate(units):
        ost, oout, oused = oracle.lzma2_decompress_raw(u, 18)
        assert st[i] == ost == 0 and outs[i] == oout == raws[i] and used[i] == oused, (i, st[i])


def test_corruption_and_check_payload(gpu, oracle):
    rng = random.Random(14)
    raw = H.textlike(20000, 53)
    xz = lzma.compress(raw, check=lzma.CHECK_CRC32)
    bad = bytearray(xz); bad[len(xz) - 12 - 8 - 4 - 2] ^= 0xFF
    ost, oout, _ = oracle.xz_unarchive(bytes(bad))
    try:
        gpu.XZArchive.unarchive(bytes(bad))
        assert ost == 0
    except gpu.SWCompressionError as e:
        assert e.code == ost
        if ost == 807:
            assert e.payload == oout
    cases = []
    for data in (xz, lzma.compress(raw, format=lzma.FORMAT_ALONE)):
        for _ in range(25):
            cases.append((data is xz, data[:rng.randrange(1, len(data))]))
        for _ in range(40):
            b = bytearray(data); b[rng.randrange(len(b))] ^= 1 << rng.randrange(8); cases.append((data is xz, bytes(b)))
    for is_xz, c in cases:
        fn, ofn = (gpu.XZArchive.unarchive, oracle.xz_unarchive) if is_xz else (gpu.LZMA.decompress, oracle.lzma_decompress)
        ost, oout, _ = ofn(c)
        try:
            out = fn(c)
            assert ost == 0 and out == oout
        except gpu.SWCompressionError as e:
            if e.code == 1:          # engine: corrupted size fields asked for more output than the engine will allocate
                continue
            assert e.code == ost, (is_xz, e.code, ost)


def test_multi_stream_and_multi_block_prefetch(gpu, oracle):
    """XZ files with many streams / many blocks: blocks are located through the stream indexes and decoded as one batch;
    the in-order parser validates everything, so results and errors equal the oracle's."""
    import subprocess
    rng = random.Random(31)
    raws = [H.textlike(rng.choice([100, 5000, 70000]), 1700 + i) for i in range(24)]
    data = b"".join(lzma.compress(r, check=rng.choice([lzma.CHECK_CRC32, lzma.CHECK_CRC64, lzma.CHECK_SHA256, lzma.CHECK_NONE])) + bytes(4 * (i % 3))
                    for i, r in enumerate(raws))
    assert gpu.XZArchive.splitUnarchive(data) == raws
    assert gpu.XZArchive.unarchive(data) == b"".join(raws)
    big = H.textlike(600000, 1800)
    mb = subprocess.run(["xz", "-c", "-T1", "--block-size=65536"], input=big, stdout=subprocess.PIPE, check=True).stdout
    assert oracle.xz_unarchive(mb)[:2] == (0, big)
    assert gpu.XZArchive.unarchive(mb) == big
    for _ in range(20):                                      # corruption anywhere: same outcome as the oracle
        b = bytearray(mb); b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
        ost, oout, _ = oracle.xz_unarchive(bytes(b))
        try:
            out = gpu.XZArchive.unarchive(bytes(b))
            assert ost == 0 and out == oout
        except gpu.SWCompressionError as e:
            if e.code == 1:
                continue
            assert e.code == ost, (e.code, ost)
            if ost == 807:
                assert e.payload == oout

"""Parity of the CUDA BZip2 path (through the C ABI) with the CPU oracle."""
import bz2
import random

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available()
    import swcompression_b200 as S
    return S


@pytest.mark.parametrize("rel,ans", H.fixtures("BZip2/"))
def test_fixtures(gpu, rel, ans):
    assert gpu.BZip2.decompress(H.fixture(rel)) == H.answer(ans)


@pytest.mark.parametrize("raw", H.ROUNDTRIP_STRINGS)
def test_roundtrip_strings(gpu, raw):
    assert gpu.BZip2.decompress(bz2.compress(raw)) == raw


def test_short_inputs(gpu, oracle):
    for n in range(0, 16):
        junk = bytes(range(n))
        ost = oracle.bzip2_decompress(junk)[0]
        with pytest.raises(gpu.SWCompressionError) as e:
            gpu.BZip2.decompress(junk)
        assert e.value.code == ost


def test_config4_shape_single_900k_block(gpu, oracle):
    raw = H.textlike(900000, 4)
    comp = bz2.compress(raw, 9)
    ost, oout, oused = oracle.bzip2_decompress(comp)
    assert ost == 0 and oout == raw
    assert gpu.BZip2.decompress(comp) == raw


def test_multi_block_and_multi_stream(gpu, oracle):
    raw = H.textlike(350000, 41)
    comp = bz2.compress(raw, 1)                      # 100 KB blocks -> 4 blocks
    assert gpu.BZip2.decompress(comp) == raw
    a, b = H.textlike(5000, 42), bytes(300000)
    assert gpu.BZip2.multiDecompress(bz2.compress(a) + bz2.compress(b)) == [a, b]
    assert gpu.BZip2.decompress(bz2.compress(a) + bz2.compress(b)) == a     # first stream only


def test_batch_ragged(oracle):
    from swcompression_b200.batch import Batch
    rng = random.Random(8)
    raws = []
    for i in range(48):
        n = rng.choice([0, 1, 4, 5, 255, 256, 1000, 50000, 200000])
        k = i % 4
        raws.append(H.textlike(max(n, 70), 900 + i)[:n] if k == 0 else bytes(n) if k == 1 else
                    bytes(rng.getrandbits(8) for _ in range(n)) if k == 2 else (b"aaaa\x00" * (n // 5 + 1))[:n])
    units = [bz2.compress(r, rng.choice([1, 9])) for r in raws]
    b = Batch.from_units("bzip2", units, 262144)
    b.run()
    st, ln, used = b.results()
    outs = b.outputs()
    for i, u in enumerate(units):
        ost, oout, oused = oracle.bzip2_decompress(u)
        assert st[i] == ost == 0 and outs[i] == oout == raws[i] and used[i] == oused, (i, st[i])


def test_truncation_corruption_and_crc_payload(gpu, oracle):
    rng = random.Random(12)
    raw = H.textlike(30000, 43)
    comp = bz2.compress(raw)
    bad = bytearray(comp); bad[10] ^= 1                                  # block CRC
    with pytest.raises(gpu.BZip2Error) as e:
        gpu.BZip2.decompress(bytes(bad))
    assert e.value.case == "wrongCRC" and e.value.payload == raw         # BZip2Tests.swift:81-97
    cases = [comp[:rng.randrange(1, len(comp))] for _ in range(40)]
    for _ in range(60):
        b = bytearray(comp); b[rng.randrange(len(b))] ^= 1 << rng.randrange(8); cases.append(bytes(b))
    for c in cases:
        ost, oout, _ = oracle.bzip2_decompress(c)
        try:
            out = gpu.BZip2.decompress(c)
            assert ost == 0 and out == oout
        except gpu.SWCompressionError as e:
            if e.code in (1, 6):         # engine: output bound exceeded by a corrupted run length / over-subscribed code set
                continue
            assert e.code == ost, (e.code, ost)
            if ost == 210:
                assert e.payload == oout


def test_multi_stream_batch_discovery(gpu, oracle):
    """BZip2.multiDecompress on many concatenated streams: streams are discovered up front and decoded as one batch;
    anything that does not validate falls back to the in-order loop, so results and errors equal the oracle's."""
    rng = random.Random(21)
    raws = [H.textlike(rng.choice([70, 500, 3000, 40000]), 1500 + i) if i % 7 else b"" for i in range(40)]
    data = b"".join(bz2.compress(r, rng.choice([1, 9])) for r in raws)
    assert gpu.BZip2.multiDecompress(data) == raws
    ost, parts, _ = oracle.bzip2_multi_decompress(data)
    assert ost == 0 and parts == raws
    # trailing garbage / a corrupted stream in the middle: same error as the sequential reference order
    for bad in (data + b"BZh9" + bytes(20), data + b"xx", data[:len(data) // 2] + b"\x00" + data[len(data) // 2 + 1:]):
        ost, parts, _ = oracle.bzip2_multi_decompress(bad)
        try:
            got = gpu.BZip2.multiDecompress(bad)
            assert ost == 0 and got == parts
        except gpu.SWCompressionError as e:
            assert e.code == ost, (e.code, ost)
    streams = [bz2.compress(r) for r in raws[:12]]
    b = bytearray(streams[5]); b[10] ^= 1; streams[5] = bytes(b)          # block CRC of stream 5
    ost, parts, _ = oracle.bzip2_multi_decompress(b"".join(streams))
    with pytest.raises(gpu.BZip2Error) as e:
        gpu.BZip2.multiDecompress(b"".join(streams))
    assert e.value.code == ost == 210 and e.value.payload == [raws[5]]


def test_block_parallel_single_stream_matches_oracle(gpu, oracle):
    """A multi-block stream is decoded block-parallel (device scan for the 48-bit magics at every bit offset, one unit per
    block, the reference's in-order walk as validator: api_bzip2.cu bzip2_stream_by_blocks); results, consumed bits, errors and
    the wrongCRC payload must equal the sequential walk of BZip2.swift:50-95."""
    rng = random.Random(19)
    raw = H.textlike(1200000, 44) + bytes(150000) + H.textlike(300000, 45)      # a run-heavy block in the middle
    for level in (1, 3, 9):
        comp = bz2.compress(raw, level)
        ost, oout, oused = oracle.bzip2_decompress(comp)
        assert ost == 0 and oout == raw
        assert gpu.BZip2.decompress(comp) == raw
        out, used = gpu.BZip2.decompress_from(comp, 0) if hasattr(gpu.BZip2, "decompress_from") else (raw, oused)
        assert out == raw and used == oused
    comp = bz2.compress(raw, 1)                                                  # 17 blocks
    cases = [comp[:rng.randrange(20, len(comp))] for _ in range(10)]
    for _ in range(30):
        b = bytearray(comp); b[rng.randrange(len(b))] ^= 1 << rng.randrange(8); cases.append(bytes(b))
    b = bytearray(comp); b[-2] ^= 0x10; cases.append(bytes(b))                   # combined CRC
    cases.append(comp + bz2.compress(b"tail"))                                   # decompress() stops after the first stream
    for c in cases:
        ost, oout, _ = oracle.bzip2_decompress(c)
        try:
            out = gpu.BZip2.decompress(c)
            assert ost == 0 and out == oout
        except gpu.SWCompressionError as e:
            if e.code in (1, 6):
                continue
            assert e.code == ost, (e.code, ost)
            if ost == 210:
                assert e.payload == oout
    # multi-block streams inside a multi-stream archive
    parts = [H.textlike(250000, 46), H.textlike(120000, 47), b"", H.textlike(330000, 48)]
    arch = b"".join(bz2.compress(p, 1) for p in parts)
    assert gpu.BZip2.multiDecompress(arch) == parts

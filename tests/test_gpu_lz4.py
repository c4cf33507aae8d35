"""Parity of the CUDA LZ4 path (raw block batches + frame layer through the C ABI) with the CPU oracle."""
import random
import struct

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


def run_blocks(units, cap):
    from swcompression_b200.batch import Batch
    b = Batch.from_units("lz4_block", units, cap)
    b.run()
    st, ln, _ = b.results()
    return st, ln, b.outputs()


@pytest.mark.parametrize("rel,ans", H.fixtures("LZ4/"))
def test_fixtures(gpu, rel, ans):                      # frames, legacy frames, B4-B7, dependent (_BD) blocks
    assert gpu.LZ4.decompress(H.fixture(rel)) == H.answer(ans)


def test_short_inputs(gpu):                            # LZ4Tests.swift:87-108
    for data, case in ((b"", "truncated"), (b"\0", "truncated"), (bytes(1 << 20), "corrupted")):
        with pytest.raises(gpu.DataError) as e:
            gpu.LZ4.decompress(data)
        assert e.value.case == case


def test_block_batch_config3_shape(oracle):
    rng = random.Random(3)
    raws = []
    for i in range(256):
        k = rng.random()
        if k < 0.1:
            raws.append(bytes(65536))
        elif k < 0.2:
            raws.append(bytes(rng.getrandbits(8) for _ in range(65536)))
        else:
            raws.append(H.textlike(65536, 3 + i))
    units = [H.lz4_block_compress(r) for r in raws]
    st, ln, outs = run_blocks(units, 65536)
    for i, u in enumerate(units):
        ost, oout, _ = oracle.lz4_block(u)
        assert st[i] == ost == 0 and outs[i] == oout == raws[i], i


def test_benched_config_two_phase_kernels_all_units(oracle):
    """BASELINE configs[2] shape through the kernels tools/bench_codecs.py times: 20 480 independent 64 KiB blocks (256 distinct:
    80 % text-like, 10 % zeros, 10 % incompressible) take lz4_parse_kernel + lz4_exec_kernel in one call.  Every distinct block is
    compared byte for byte with the oracle, and every tiled copy with the first copy on the device."""
    import numpy as np
    import torch  # noqa: F401
    from swcompression_b200.batch import Batch, pack_units
    rng = random.Random(33)
    raws = []
    for i in range(256):
        k = rng.random()
        raws.append(bytes(65536) if k < 0.1 else bytes(rng.getrandbits(8) for _ in range(65536)) if k < 0.2 else H.textlike(65536, 700 + i))
    units = [H.lz4_block_compress(r) for r in raws]
    buf, offs, lens = pack_units(units)
    stride, tile = len(buf) - 64, 80
    big = np.concatenate([np.tile(buf[:stride], tile), np.zeros(64, dtype=np.uint8)])
    all_off = (offs[None, :] + (np.arange(tile, dtype=np.uint64) * np.uint64(stride))[:, None]).reshape(-1)
    b = Batch("lz4_block", big, all_off, np.tile(lens, tile), 65536)
    assert b.n == 20480
    b.run()
    st, ln, _ = b.results()
    assert (st == 0).all() and (ln == 65536).all()
    out = b.d_out[: b.n * 65536].view(tile, 256, 65536)
    assert bool((out == out[0:1]).all()), "tiled copies differ"
    first = out[0].cpu().numpy()
    for i, u in enumerate(units):
        ost, oout, _ = oracle.lz4_block(u)
        assert ost == 0 and bytes(first[i]) == oout == raws[i], i


def test_block_ragged_and_fuzz(oracle):
    rng = random.Random(4)
    units = []
    for i in range(150):
        n = rng.choice([1, 5, 12, 13, 64, 300, 4000, 70000])
        raw = H.textlike(max(n, 70), 200 + i)[:n] if i % 3 else bytes(n)
        units.append(H.lz4_block_compress(raw))
    good = list(units)
    for u in good[:80]:
        if len(u) > 2:
            units.append(u[:rng.randrange(1, len(u))])
            b = bytearray(u); b[rng.randrange(len(b))] ^= 1 << rng.randrange(8); units.append(bytes(b))
    units.append(b"")
    st, ln, outs = run_blocks(units, 1 << 18)
    for i, u in enumerate(units):
        ost, oout, _ = oracle.lz4_block(u)
        assert st[i] == ost, (i, st[i], ost)
        if ost == 0:
            assert outs[i] == oout


def test_block_overflow(oracle):
    raws = [H.textlike(20000, 300 + i) for i in range(8)]
    units = [H.lz4_block_compress(r) for r in raws]
    st, ln, outs = run_blocks(units, 10000)
    assert (st == 1).all() and (ln == 20000).all()


def test_frames_roundtrip_and_errors(gpu, oracle):
    rng = random.Random(6)
    raws = [H.textlike(65536, 400 + i) for i in range(5)] + [H.textlike(1234, 77)]
    for cck in (False, True):
        for bck in (False, True):
            f = H.lz4_frame_independent(raws, content_checksum=cck, block_checksum=bck)
            assert gpu.LZ4.decompress(f) == b"".join(raws)
            for _ in range(6):
                cut = f[:rng.randrange(1, len(f))]
                ost = oracle.lz4_decompress(cut)[0]
                with pytest.raises(gpu.SWCompressionError) as e:
                    gpu.LZ4.decompress(cut)
                assert e.value.code == ost
            for _ in range(6):
                b = bytearray(f); b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
                ost, oout, _ = oracle.lz4_decompress(bytes(b))
                try:
                    out = gpu.LZ4.decompress(bytes(b))
                    assert ost == 0 and out == oout
                except gpu.SWCompressionError as e:
                    assert e.code == ost
                    if ost == 503:
                        assert e.payload == oout
    f1 = H.lz4_frame_independent(raws[:2]); f2 = H.lz4_frame_independent(raws[2:4], content_checksum=True)
    skip = struct.pack("<II", 0x184D2A53, 5) + b"hello"
    assert gpu.LZ4.multiDecompress(f1 + skip + f2) == [b"".join(raws[:2]), b"".join(raws[2:4])]
    assert gpu.LZ4.decompress(skip + f2) == b"".join(raws[2:4])


def test_dictionary_independent_blocks(gpu, oracle):
    dic = H.textlike(4096, 500)
    raw = dic[1000:3000] + H.textlike(3000, 501)
    # hand-built block: 4 literals, then a match reaching 2000 bytes back into the dictionary, then 8 literals
    blk = bytes([0x4F]) + raw[:4] + struct.pack("<H", 2000 + 4) + bytes([100]) + bytes([0x80]) + raw[4:12]
    desc = bytes([0x60, 0x40])
    import oracle_xxh
    frame = struct.pack("<I", 0x184D2204) + desc + bytes([(oracle_xxh.xxh32(desc) >> 8) & 0xFF]) + struct.pack("<I", len(blk)) + blk + struct.pack("<I", 0)
    ost, oout, _ = oracle.lz4_decompress(frame, dictionary=dic)
    assert ost == 0
    assert gpu.LZ4.decompress(frame, dictionary=dic) == oout
    ost2 = oracle.lz4_decompress(frame)[0]
    with pytest.raises(gpu.DataError) as e:
        gpu.LZ4.decompress(frame)
    assert e.value.code == ost2 == 502

"""The parts of the C ABI added for the drop-in boundary: raw-LZMA batches, batched checksum epilogues,
concurrent host threads, and hostile XZ indexes."""
import ctypes as C
import lzma
import random
import struct
import threading
import zlib

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


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64) if a.dtype == np.uint64 else np.ascontiguousarray(a)).cuda()


def test_lzma_raw_batch_matches_oracle(gpu, oracle):
    """swc_lzma_decompress_batch: one raw LZMA stream per unit with its own properties / dictionary size / size (the ZIP and
    7-Zip form, LZMA.swift:56-61) == the oracle's LZMA.decompress(data:properties:uncompressedSize:) unit by unit."""
    import torch
    from swcompression_b200 import _lib
    from swcompression_b200.batch import pack_units
    rng = random.Random(3)
    units, props, dsz, usz, raws = [], [], [], [], []
    for i in range(48):
        lc, lp, pb = rng.choice([(3, 0, 2), (0, 2, 1), (2, 2, 0), (4, 0, 4), (1, 3, 3), (0, 0, 0)])
        d = rng.choice([1 << 12, 1 << 16, 1 << 20])
        raw = H.textlike(rng.randrange(1, 40000), 7000 + i) if i % 7 else b""
        alone = lzma.compress(raw, format=lzma.FORMAT_ALONE, filters=[{"id": lzma.FILTER_LZMA1, "lc": lc, "lp": lp, "pb": pb, "dict_size": d}])
        body = alone[13:]                                # .lzma = 13-byte header + raw stream with end marker
        known = i % 2 == 0
        units.append(body); props.append(lc | lp << 8 | pb << 16); dsz.append(d); usz.append(len(raw) if known else -1); raws.append(raw)
    # damaged and truncated units ride along
    for i in (3, 8, 13):
        b = bytearray(units[i]); b[len(b) // 2] ^= 0x55; units.append(bytes(b))
        props.append(props[i]); dsz.append(dsz[i]); usz.append(usz[i]); raws.append(None)
        units.append(units[i][: max(len(units[i]) // 2, 1)]); props.append(props[i]); dsz.append(dsz[i]); usz.append(usz[i]); raws.append(None)
    n = len(units)
    buf, offs, lens = pack_units(units)
    cap = 40960
    o_off = np.arange(n, dtype=np.uint64) * np.uint64(cap)
    d_in, d_off, d_len = _dev(buf), _dev(offs), _dev(lens)
    d_props, d_dsz, d_usz = _dev(np.array(props, dtype=np.uint32)), _dev(np.array(dsz, dtype=np.int64)), _dev(np.array(usz, dtype=np.int64))
    d_ooff, d_ocap = _dev(o_off), _dev(np.full(n, cap, dtype=np.uint64))
    d_out = torch.zeros(n * cap + 64, dtype=torch.uint8, device="cuda")
    d_olen = torch.zeros(n, dtype=torch.int64, device="cuda"); d_used = torch.zeros(n, dtype=torch.int64, device="cuda")
    d_st = torch.full((n,), -1, dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = _lib.lib().swc_lzma_decompress_batch(p(d_in), p(d_off), p(d_len), p(d_props), p(d_dsz), p(d_usz), p(d_out), p(d_ooff), p(d_ocap),
                                              p(d_olen), p(d_used), p(d_st), n, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    torch.cuda.synchronize()
    st, ln, used, host = d_st.cpu().numpy(), d_olen.cpu().numpy(), d_used.cpu().numpy(), d_out.cpu().numpy()
    ok = 0
    for i in range(n):
        lc, lp, pb = props[i] & 255, (props[i] >> 8) & 255, props[i] >> 16
        ost, oout, oused = oracle.lzma_decompress_raw(units[i], lc, lp, pb, dsz[i], None if usz[i] < 0 else usz[i])
        assert st[i] == ost, (i, st[i], ost)
        if ost == 0:
            ok += 1
            assert bytes(host[i * cap:i * cap + ln[i]]) == oout and used[i] == oused
            if raws[i] is not None:
                assert oout == raws[i]
    assert ok >= 48


def test_checksum_batches(gpu, oracle):
    import torch
    from swcompression_b200 import _lib
    from swcompression_b200.batch import pack_units
    rng = random.Random(4)
    units = [bytes(rng.getrandbits(8) for _ in range(rng.choice([0, 1, 3, 4, 15, 16, 17, 1000, 65536, 100001]))) for _ in range(40)]
    buf, offs, lens = pack_units(units)
    d_in, d_off, d_len = _dev(buf), _dev(offs), _dev(lens)
    st = np.zeros(len(units), dtype=np.int32); st[5] = 7                     # a unit that "did not decode" is not read: result 0
    d_st = _dev(st)
    d_crc = torch.full((len(units),), 0x0BADF00D, dtype=torch.int32, device="cuda")
    d_xxh = torch.zeros(len(units), dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    L = _lib.lib()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    assert L.swc_crc32_batch(p(d_in), p(d_off), p(d_len), p(d_st), p(d_crc), len(units), s) == 0
    assert L.swc_xxh32_batch(p(d_in), p(d_off), p(d_len), p(d_xxh), len(units), s) == 0
    torch.cuda.synchronize()
    crc = d_crc.cpu().numpy().view(np.uint32); xxh = d_xxh.cpu().numpy().view(np.uint32)
    for i, u in enumerate(units):
        if i == 5:
            assert crc[i] == 0
        else:
            assert crc[i] == zlib.crc32(u) == oracle.lib().swco_crc32(u, len(u), 0)
        assert xxh[i] == oracle.lib().swco_xxh32(u, len(u))


def test_concurrent_host_threads(gpu, oracle):
    """The reference is re-entrant; the library serialises per device behind a mutex (include/swcgpu.h "threading").  Four
    threads hammer different entry points at once (ctypes drops the GIL inside the calls) and every result must be right."""
    import bz2
    import gzip
    raws = [H.textlike(20000 + 1000 * i, 8000 + i) for i in range(6)]
    jobs = {
        "deflate": [(H.raw_deflate(r), r) for r in raws],
        "gzip": [(gzip.compress(r), r) for r in raws],
        "bzip2": [(bz2.compress(r), r) for r in raws],
        "xz": [(lzma.compress(r), r) for r in raws],
        "lz4": [(H.lz4_frame_independent([r], content_checksum=True), r) for r in raws],
        "zlib": [(zlib.compress(r), r) for r in raws],
    }
    fn = {"deflate": gpu.Deflate.decompress, "gzip": gpu.GzipArchive.unarchive, "bzip2": gpu.BZip2.decompress,
          "xz": gpu.XZArchive.unarchive, "lz4": gpu.LZ4.decompress, "zlib": gpu.ZlibArchive.unarchive}
    errors = []

    def worker(seed):
        rng = random.Random(seed)
        try:
            for _ in range(40):
                k = rng.choice(list(jobs))
                data, want = rng.choice(jobs[k])
                got = fn[k](data)
                if got != want:
                    errors.append((seed, k))
                if rng.random() < 0.2 and gpu.crc32(want) != zlib.crc32(want):
                    errors.append((seed, "crc32"))
        except Exception as e:            # noqa: BLE001
            errors.append((seed, repr(e)))

    ts = [threading.Thread(target=worker, args=(s,)) for s in range(4)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors[:5]


def _vli(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def test_xz_hostile_index_is_an_error_not_a_crash(gpu, oracle):
    """index records are untrusted.  Streams whose index claims block sizes up to 2^63-1 (sums that wrap 64
    bits), more records than bytes, or sizes reaching before the stream header must come back with the oracle's status."""
    base = lzma.compress(H.textlike(5000, 1), check=lzma.CHECK_CRC32)
    two = lzma.compress(H.textlike(5000, 1), check=lzma.CHECK_CRC32) + lzma.compress(H.textlike(3000, 2), check=lzma.CHECK_CRC32)
    header, footer_flags = base[:12], base[-4:-2]

    def stream(records, blocks=b""):
        idx = b"\x00" + _vli(len(records)) + b"".join(_vli(a) + _vli(b) for a, b in records)
        idx += b"\x00" * (-len(idx) % 4)
        idx += struct.pack("<I", zlib.crc32(idx))
        backward = len(idx) // 4 - 1
        ft = struct.pack("<I", backward) + footer_flags
        ft = struct.pack("<I", zlib.crc32(ft)) + ft + b"YZ"
        return header + blocks + idx + ft

    big = (1 << 63) - 1
    cases = [
        stream([(big, 1), (big, 1)]),                                   # padded sizes sum to 0 mod 2^64
        stream([(big, 1)] * 4),
        stream([(1 << 62, 5), (1 << 62, 5), (1 << 62, 5), (1 << 62, 5)]),
        stream([(100, 100)] * 3),                                       # sizes reach before the stream header
        stream([(8, 1 << 40), (8, 1 << 40)], blocks=b"\x00" * 16),      # absurd uncompressed sizes
        stream([(5, 1)], blocks=b"\x02\x00\x21\x01\x00\x00\x00\x00"),
        two[:-12] + stream([(big, 7), (big, 7)])[12:],                  # hostile index behind a healthy first stream
    ]
    rng = random.Random(9)
    for _ in range(150):                                                # byte flips in index / footer of a two-stream archive
        b = bytearray(two)
        for _ in range(rng.randrange(1, 4)):
            b[len(b) - 1 - rng.randrange(40)] ^= 1 << rng.randrange(8)
        cases.append(bytes(b))
    for c in cases:
        ost, oparts, _ = oracle.xz_split_unarchive(c)
        if ost == 0:
            assert gpu.XZArchive.splitUnarchive(c) == oparts
        else:
            with pytest.raises(gpu.SWCompressionError) as e:
                gpu.XZArchive.splitUnarchive(c)
            assert e.value.code == ost, (e.value.code, ost)

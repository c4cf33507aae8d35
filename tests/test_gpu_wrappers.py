"""Parity of the GZip / Zlib wrappers and the device checksums with the oracle and the golden fixtures."""
import gzip
import hashlib
import random
import zlib

import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import torch
    assert torch.cuda.is_available()
    import swcompression_b200 as S
    return S


@pytest.mark.parametrize("rel,ans", H.fixtures("GZip/"))
def test_gzip_fixtures(gpu, rel, ans):
    assert gpu.GzipArchive.unarchive(H.fixture(rel)) == H.answer(ans)


def test_zlib_fixtures(gpu):
    assert gpu.ZlibArchive.unarchive(H.fixture("Zlib/test_empty.zlib")) == b""
    with pytest.raises(gpu.DeflateError) as e:
        gpu.ZlibArchive.unarchive(H.fixture("Zlib/test.zlib"))
    assert e.value.case == "wrongBlockType"


def test_checksums(gpu, oracle):
    rng = random.Random(1)
    for n in (0, 1, 15, 16, 17, 4095, 4096, 4097, 70001, 1 << 20):
        b = bytes(rng.getrandbits(8) for _ in range(min(n, 70001))) * (1 if n <= 70001 else 15)
        b = b[:n]
        assert gpu.crc32(b) == zlib.crc32(b) == oracle.crc32(b)
        assert gpu.adler32(b) == zlib.adler32(b)
        assert gpu.crc64(b) == oracle.crc64(b)
        assert gpu.bzip2_crc32(b) == oracle.bzip2_crc32(b)
        assert gpu.xxh32(b) == oracle.xxh32(b)
        if n <= 70001:
            assert gpu.sha256(b) == hashlib.sha256(b).digest()
    from test_oracle_golden import XXH
    for msg, h in XXH:
        assert gpu.xxh32(msg) == h


def test_gzip_zlib_roundtrip_errors_and_payloads(gpu, oracle):
    rng = random.Random(2)
    raw = H.textlike(50000, 21)
    assert gpu.GzipArchive.unarchive(gzip.compress(raw)) == raw
    assert gpu.ZlibArchive.unarchive(zlib.compress(raw)) == raw
    a, b = H.textlike(5000, 22), H.textlike(70000, 23)
    assert gpu.GzipArchive.multiUnarchive(gzip.compress(a) + gzip.compress(b)) == [a, b]
    g = bytearray(gzip.compress(raw)); g[-8] ^= 1
    with pytest.raises(gpu.GzipError) as e:
        gpu.GzipArchive.unarchive(bytes(g))
    assert e.value.case == "wrongCRC" and e.value.payload == raw
    z = bytearray(zlib.compress(raw)); z[-1] ^= 1
    with pytest.raises(gpu.ZlibError) as e:
        gpu.ZlibArchive.unarchive(bytes(z))
    assert e.value.case == "wrongAdler32" and e.value.payload == raw
    for make, fn, ofn in ((gzip.compress, gpu.GzipArchive.unarchive, oracle.gzip_unarchive),
                          (zlib.compress, gpu.ZlibArchive.unarchive, oracle.zlib_unarchive)):
        data = make(raw)
        for _ in range(12):
            cut = data[:rng.randrange(1, len(data))]
            ost = ofn(cut)[0]
            with pytest.raises(gpu.SWCompressionError) as e:
                fn(cut)
            assert e.value.code == ost

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
    assert [m.data for m in gpu.GzipArchive.multiUnarchive(gzip.compress(a) + gzip.compress(b))] == [a, b]
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


def _members(rng, count):
    """count gzip members of mixed size/shape: empty, named, stored payloads holding the member signature, text."""
    raws, blobs = [], []
    for i in range(count):
        kind = i % 7
        if kind == 0:
            raw = b""
        elif kind == 1:
            raw = (b"\x1f\x8b\x08\x00" + bytes(rng.randrange(256) for _ in range(40))) * 30   # look-alike signatures
        else:
            raw = H.textlike(rng.randrange(1, 70000), 100 + i)
        level = 0 if kind == 1 else rng.choice((1, 6, 9))
        blob = gzip.compress(raw, level)
        if kind == 3:                                  # FNAME header
            blob = blob[:3] + b"\x08" + blob[4:10] + b"member-%d\x00" % i + blob[10:]
        raws.append(raw)
        blobs.append(blob)
    return raws, blobs


def test_gzip_multi_member_batch_matches_oracle(gpu, oracle):
    """Many-member archives take the batched path (signature scan + speculative decode + in-order validation); the result,
    the error case and the members returned with it must equal the sequential reference walk (GzipArchive.swift:52-77)."""
    rng = random.Random(77)
    raws, blobs = _members(rng, 60)
    data = b"".join(blobs)
    ost, oparts, _ = oracle.gzip_multi_unarchive(data)
    assert ost == 0 and oparts == raws
    assert [m.data for m in gpu.GzipArchive.multiUnarchive(data)] == raws
    # wrong CRC in member 17: members 0..17 are returned with the error
    bad = list(blobs)
    b = bytearray(bad[17]); b[-8] ^= 0x40; bad[17] = bytes(b)
    data = b"".join(bad)
    ost, oparts, _ = oracle.gzip_multi_unarchive(data)
    with pytest.raises(gpu.GzipError) as e:
        gpu.GzipArchive.multiUnarchive(data)
    assert e.value.code == ost and e.value.case == "wrongCRC" and [m.data for m in e.value.payload] == oparts == raws[:18]
    # wrong ISIZE in member 30, damaged payload in member 5, trailing garbage, truncation: same error as the oracle
    variants = []
    b = bytearray(blobs[30]); b[-1] ^= 1
    variants.append(b"".join(blobs[:30]) + bytes(b) + b"".join(blobs[31:]))
    b = bytearray(blobs[5]); b[len(b) // 2] ^= 0xFF
    variants.append(b"".join(blobs[:5]) + bytes(b) + b"".join(blobs[6:]))
    variants.append(b"".join(blobs) + b"\x00" * 7)
    variants.append(b"".join(blobs) + b"\x1f\x8b\x08\x00" + b"\x00" * 30)
    variants.append(b"".join(blobs)[:-3])
    for v in variants:
        ost, oparts, _ = oracle.gzip_multi_unarchive(v)
        if ost == 0:
            assert [m.data for m in gpu.GzipArchive.multiUnarchive(v)] == oparts
        else:
            with pytest.raises(gpu.SWCompressionError) as e:
                gpu.GzipArchive.multiUnarchive(v)
            assert e.value.code == ost
    # BGZF-shaped archive: 3000 members of <= 64 KiB (thread-per-member kernel would need >= 20000; this is the warp kernel)
    pool = [gzip.compress(H.textlike(rng.randrange(20000, 65536), 900 + i), 6) for i in range(24)]
    praw = [gzip.decompress(p) for p in pool]
    order = [rng.randrange(24) for _ in range(3000)]
    got = gpu.GzipArchive.multiUnarchive(b"".join(pool[i] for i in order))
    assert len(got) == 3000 and all(got[j].data == praw[order[j]] for j in range(3000))


def test_gzip_members_carry_their_headers(gpu):
    """GzipArchive.multiUnarchive -> [Member] (GzipArchive.swift:13-22,52-77): every member comes with the GzipHeader parsed at
    its own offset inside the archive — sequential walk (few members) and batched walk (many members) alike."""
    import io
    def member(raw, name, mtime, comment=None):
        bio = io.BytesIO()
        with gzip.GzipFile(filename=name, mode="wb", fileobj=bio, mtime=mtime) as f:
            f.write(raw)
        blob = bytearray(bio.getvalue())
        if comment is not None:                  # add FCOMMENT by hand (python's gzip never writes one)
            assert blob[3] == 0x08
            name_end = blob.index(0, 10) + 1
            blob[3] |= 0x10
            blob[name_end:name_end] = comment.encode("latin-1") + b"\0"
        return bytes(blob)
    for count in (3, 40):
        raws = [H.textlike(300 + 17 * i, 4000 + i) for i in range(count)]
        blobs = [member(raws[i], f"file_{i}.txt", 1500000000 + i, comment=(f"c{i}" if i % 3 == 0 else None)) for i in range(count)]
        ms = gpu.GzipArchive.multiUnarchive(b"".join(blobs))
        assert [m.data for m in ms] == raws
        for i, m in enumerate(ms):
            assert m.header.fileName == f"file_{i}.txt" and m.header.compressionMethod == "deflate"
            assert int(m.header.modificationTime.timestamp()) == 1500000000 + i
            assert m.header.comment == (f"c{i}" if i % 3 == 0 else None)
            assert m.header.extraFields == []
    h = gpu.GzipHeader(H.fixture("GZip/test1.gz"))               # GzipTests.swift header checks: name + mtime of test1.gz
    assert h.fileName == "test1.answer" and int(h.modificationTime.timestamp()) == 1482698300 and h.osType == "unix"
    assert not h.isTextFile and h.comment is None
    z = gpu.ZlibHeader(zlib.compress(b"abc", 9))
    assert z.windowSize == 32768 and z.compressionLevel == "slowAlgorithm" and z.compressionMethod == "deflate"

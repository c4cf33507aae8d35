"""ZIP container: ZipContainer.open / info through the C ABI (swc_zip_open / swc_zip_info) vs the oracle's
restatement of Sources/ZIP/*.swift — and the oracle itself vs Python's zipfile on archives zipfile writes (stored, deflate,
bzip2, lzma, directories, data descriptors, ZIP64, comments) plus hand-made damaged containers whose expected error is
read off the Swift source."""
import io
import random
import struct
import zipfile

import pytest

import helpers as H

METHODS = [zipfile.ZIP_STORED, zipfile.ZIP_DEFLATED, zipfile.ZIP_BZIP2, zipfile.ZIP_LZMA]


class _NoSeek(io.RawIOBase):
    """an unseekable sink makes zipfile write data descriptors (general purpose bit 3)"""

    def __init__(self):
        self.b = bytearray()

    def writable(self):
        return True

    def write(self, d):
        self.b += d
        return len(d)


def build(entries, streamed=False, zip64=False, comment=b""):
    """entries: list of (name, bytes or None for a directory, method)"""
    sink = _NoSeek() if streamed else io.BytesIO()
    with zipfile.ZipFile(sink, "w") as z:
        for name, data, method in entries:
            if data is None:
                z.mkdir(name)
            else:
                zi = zipfile.ZipInfo(name, date_time=(2021, 3, 4, 5, 6, 8))
                zi.compress_type = method
                with z.open(zi, "w", force_zip64=zip64) as f:
                    f.write(data)
        z.comment = comment
    return bytes(sink.b) if streamed else sink.getvalue()


def corpus(seed, count):
    rng = random.Random(seed)
    out = []
    for i in range(count):
        kind = rng.randrange(5)
        n = rng.choice([0, 1, 10, 300, 5000, 70000])
        raw = (H.textlike(max(n, 70), seed * 100 + i)[:n] if kind < 3 else bytes(rng.getrandbits(8) for _ in range(n)) if kind == 3 else bytes(n))
        out.append((f"dir{i % 3}/file_{i}.bin", raw, METHODS[i % 4]))
    out.insert(1, ("dir0", None, 0))
    return out


def archives():
    yield "mixed", build(corpus(1, 14))
    yield "data descriptors", build(corpus(2, 9), streamed=True)
    yield "zip64 local fields", build(corpus(3, 7), zip64=True)
    yield "zip64 + descriptors", build(corpus(4, 6), streamed=True, zip64=True)
    yield "comment + utf8 names", build([("café/über.txt", b"hello", zipfile.ZIP_DEFLATED), ("plain.txt", b"", zipfile.ZIP_STORED)], comment=b"archive comment")
    yield "empty", build([])
    yield "many deflate entries", build([(f"e{i}", H.textlike(2000 + i, 9000 + i), zipfile.ZIP_DEFLATED) for i in range(300)])


def expect_from_zipfile(blob):
    with zipfile.ZipFile(io.BytesIO(blob)) as z:
        return [(i.filename, None if i.is_dir() else z.read(i), i.CRC, i.file_size, i.compress_type) for i in z.infolist()]


@pytest.mark.parametrize("name,blob", list(archives()), ids=[a[0] for a in archives()])
def test_oracle_matches_zipfile(oracle, name, blob):
    st, ents = oracle.zip_open(blob)
    assert st == 0
    want = expect_from_zipfile(blob)
    assert len(ents) == len(want)
    for e, (fn, data, crc, size, method) in zip(ents, want):
        assert e["name"].decode("utf-8" if e["utf8"] else "cp437") == fn
        assert e["is_directory"] == (data is None) and e["data"] == data
        if data is not None:
            assert e["crc"] == crc and e["size"] == size and e["method"] == method


def _patch(blob, where, fmt, value):
    b = bytearray(blob)
    struct.pack_into(fmt, b, where, value)
    return bytes(b)


def damaged():
    """(name, container, expected status) — statuses derived by hand from the Swift source"""
    base = build([("a.txt", H.textlike(3000, 5), zipfile.ZIP_DEFLATED), ("b.txt", b"stored bytes", zipfile.ZIP_STORED)])
    eocd = base.rindex(b"PK\x05\x06")
    cd = struct.unpack_from("<I", base, eocd + 16)[0]
    lh2 = struct.unpack_from("<I", base, base.index(b"PK\x01\x02", cd + 4) + 42)[0]
    yield "too short", b"PK\x05\x06" + bytes(10), 901                                   # ZipContainer.swift:139 (< 22 bytes)
    yield "no end record", bytes(100), 901                                               # :153
    yield "cd signature", _patch(base, cd, "<I", 0x02014b51), 902                        # ZipCentralDirectoryEntry.swift:47
    yield "local signature", _patch(base, 0, "<I", 0x04034b51), 902                      # ZipLocalHeader.swift:41
    yield "local version needed 64", _patch(base, 4, "<H", 64), 904                      # ZipLocalHeader.swift:117
    yield "cd version needed 0x0140 (low byte 64)", _patch(base, cd + 6, "<H", 0x0140), 904       # :128
    yield "encrypted (both headers)", _patch(_patch(base, 6, "<H", 1), cd + 8, "<H", 1), 906      # :119-122
    yield "strong encryption bit 6", _patch(_patch(base, 6, "<H", 0x40), cd + 8, "<H", 0x40), 906
    yield "patched data bit 5", _patch(_patch(base, 6, "<H", 0x20), cd + 8, "<H", 0x20), 907      # :123
    yield "flags differ between headers", _patch(base, cd + 8, "<H", 0x800), 909         # :134
    yield "method differs", _patch(base, 8, "<H", 0), 909
    yield "mod time differs", _patch(base, 10, "<H", 1), 909
    yield "disk numbers differ", _patch(base, eocd + 4, "<H", 1), 905                    # ZipEndOfCentralDirectory.swift:27
    yield "entry counts differ", _patch(base, eocd + 8, "<H", 7), 905                    # :33
    yield "entry on another disk", _patch(base, cd + 34, "<H", 3), 905                   # ZipLocalHeader.swift:130
    yield "method 99 in both headers", _patch(_patch(base, 8, "<H", 99), cd + 10, "<H", 99), 908   # ZipContainer.swift:92 (at open time)
    yield "crc of entry 2", _patch(base, lh2 + 14, "<I", 0xDEADBEEF), 910
    yield "uncompressed size of entry 1", _patch(base, 22, "<I", 2999), 903              # ZipContainer.swift:114
    yield "compressed size of entry 1", _patch(base, 18, "<I", 7), 903
    yield "name is not UTF-8 although flagged", _patch(_patch(_patch(base, 6, "<H", 0x800), cd + 8, "<H", 0x800), cd + 46, "<B", 0xFF), 911
    yield "central directory offset past the end", _patch(base, eocd + 16, "<I", len(base) + 5), 2     # unguarded read: trap


@pytest.mark.parametrize("name,blob,status", list(damaged()), ids=[d[0] for d in damaged()])
def test_damaged_containers_oracle_and_info(oracle, name, blob, status):
    import swcompression_b200 as S
    ost, ents = oracle.zip_open(blob)
    assert ost == status, (name, ost)
    # swc_zip_info is pure host code: same verdict for everything that fails before entry data is touched
    ist, _ = oracle.zip_open(blob, info_only=True)
    try:
        S.ZipContainer.info(blob)
        got = 0
    except S.SWCompressionError as e:
        got = e.code
    assert got == ist, (name, got, ist)


@pytest.mark.gpu
@pytest.mark.parametrize("name,blob", list(archives()), ids=[a[0] for a in archives()])
def test_gpu_open_matches_oracle(oracle, name, blob):
    import swcompression_b200 as S
    st, ents = oracle.zip_open(blob)
    got = S.ZipContainer.open(blob)
    assert st == 0 and len(got) == len(ents)
    for g, e in zip(got, ents):
        assert g.info.name.encode("utf-8" if e["utf8"] else "cp437") == e["name"]
        assert (g.info.type == "directory") == e["is_directory"] and g.data == e["data"]
        assert g.info.crc == e["crc"] and g.info.size == e["size"]
    infos = S.ZipContainer.info(blob)
    assert [i.name for i in infos] == [g.info.name for g in got]


@pytest.mark.gpu
def test_gpu_damaged_and_fuzzed_containers_match_oracle(oracle):
    import swcompression_b200 as S
    cases = [b for _, b, _ in damaged()]
    rng = random.Random(21)
    for _, blob in list(archives())[:5]:
        for _ in range(60):
            b = bytearray(blob)
            for _ in range(rng.randrange(1, 4)):
                b[rng.randrange(len(b))] ^= 1 << rng.randrange(8)
            cases.append(bytes(b))
        for _ in range(10):
            cases.append(blob[: rng.randrange(1, len(blob))])
    n_crc = 0
    for c in cases:
        ost, oents = oracle.zip_open(c)
        try:
            got = S.ZipContainer.open(c)
            gst = 0
        except S.SWCompressionError as e:
            gst, got = e.code, e.payload
        assert gst == ost, (gst, ost)
        if ost in (0, 910):
            n_crc += ost == 910
            assert [g.data for g in got] == [e["data"] for e in oents]
    assert n_crc > 0


@pytest.mark.gpu
def test_gpu_large_container_is_one_batch_per_method():
    """3 000 Deflate entries + BZip2 / LZMA / stored ones: the whole container costs a handful of kernel launches, not one per entry."""
    import swcompression_b200 as S
    from swcompression_b200 import _lib
    ents = [(f"d/{i}.txt", H.textlike(3000 + (i % 50) * 100, 20000 + i % 200), zipfile.ZIP_DEFLATED) for i in range(3000)]
    ents += [(f"b/{i}", H.textlike(20000, 30000 + i), zipfile.ZIP_BZIP2) for i in range(20)]
    ents += [(f"l/{i}", H.textlike(20000, 31000 + i), zipfile.ZIP_LZMA) for i in range(20)]
    ents += [(f"s/{i}", H.textlike(500, 32000 + i), zipfile.ZIP_STORED) for i in range(50)]
    blob = build(ents)
    S.ZipContainer.open(build(ents[:3]))                                  # warm-up (module load, arenas)
    before = _lib.lib().swc_kernel_launches()
    got = S.ZipContainer.open(blob)
    launches = _lib.lib().swc_kernel_launches() - before
    assert [g.data for g in got] == [e[1] for e in ents]
    assert launches < 40, launches

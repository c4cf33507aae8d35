"""XZ container and LZ4 frame framing: hand-built streams with the status the reference produces, read off the Swift source
(Sources/XZ/XZArchive.swift:27-218, XZBlock.swift:18-97, XZStreamHeader.swift:33-57, LittleEndianByteReader+XZ.swift:10-30,
Sources/LZ4/LZ4.swift:159-331) — NOT generated with the oracle.  The same table is applied to the oracle (CPU) and to the
product through the C ABI (GPU), so the two framing implementations are pinned independently of each other."""
import lzma
import struct
import zlib

import pytest

import helpers as H
from oracle_xxh import xxh32

OK = 0
XZ_MAGIC, XZ_FIELD, XZ_INFO_CRC, XZ_FILTER, XZ_DATA_SIZE, XZ_CHECK, XZ_PADDING, XZ_VLI = 801, 802, 803, 804, 806, 807, 808, 809
LZMA2_DICT = 401
TRUNC, CORRUPT, MISMATCH = 501, 502, 503


def vli(v):
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def crc(b):
    return struct.pack("<I", zlib.crc32(b))


RAW = H.textlike(3000, 77)
LZMA2 = lzma.compress(RAW, format=lzma.FORMAT_RAW, filters=[{"id": lzma.FILTER_LZMA2, "preset": 1, "dict_size": 1 << 16}])
CHECKS = {0: b"", 1: struct.pack("<I", zlib.crc32(RAW))}


def xz_header(check=1, flags0=0, bad_crc=False, magic=b"\xFD7zXZ\x00"):
    fl = bytes([flags0, check])
    c = crc(fl)
    if bad_crc:
        c = bytes([c[0] ^ 1]) + c[1:]
    return magic + fl + c


def xz_block(check=1, flags=0, comp=None, uncomp=None, filt=b"\x21\x01\x10", hdr_pad_byte=0, bad_hdr_crc=False, data=LZMA2,
             block_pad_byte=0, check_bytes=None):
    body = bytes([flags | (0x40 if comp is not None else 0) | (0x80 if uncomp is not None else 0)])
    if comp is not None:
        body += comp if isinstance(comp, bytes) else vli(comp)
    if uncomp is not None:
        body += vli(uncomp)
    body += filt
    size = (1 + len(body) + 4 + 3) // 4 * 4
    pad = size - 4 - 1 - len(body)
    hdr = bytes([size // 4 - 1]) + body + bytes([hdr_pad_byte]) * pad
    c = crc(hdr)
    if bad_hdr_crc:
        c = bytes([c[0] ^ 1]) + c[1:]
    unpadded = len(hdr) + 4 + len(data)
    blk = hdr + c + data + bytes([block_pad_byte]) * (-unpadded % 4)
    ck = CHECKS[check] if check_bytes is None else check_bytes
    return blk + ck, unpadded + len(ck)


def xz_index(records, count=None, pad_byte=0, bad_crc=False):
    idx = b"\x00" + vli(len(records) if count is None else count) + b"".join(vli(a) + vli(b) for a, b in records)
    idx += bytes([pad_byte]) * (-len(idx) % 4)
    c = crc(idx)
    if bad_crc:
        c = bytes([c[0] ^ 1]) + c[1:]
    return idx + c


def xz_footer(index_len, check=1, backward=None, flags=None, bad_crc=False, magic=b"YZ"):
    bw = struct.pack("<I", (index_len // 4 - 1) if backward is None else backward)
    fl = bytes([0, check]) if flags is None else flags
    c = crc(bw + fl)
    if bad_crc:
        c = bytes([c[0] ^ 1]) + c[1:]
    return c + bw + fl + magic


def xz_stream(check=1, header=None, block_kw=None, index_kw=None, footer_kw=None, records=None):
    blk, unpadded = xz_block(check=check, **(block_kw or {}))
    idx = xz_index(records if records is not None else [(unpadded, len(RAW))], **(index_kw or {}))
    return (header if header is not None else xz_header(check)) + blk + idx + xz_footer(len(idx), check, **(footer_kw or {}))


GOOD = xz_stream()


def xz_cases():
    yield "good crc32", GOOD, OK
    yield "good, no check", xz_stream(check=0), OK
    yield "good, sizes in the block header", xz_stream(block_kw=dict(comp=len(LZMA2), uncomp=len(RAW))), OK
    yield "shorter than 32 bytes", GOOD[:31], XZ_MAGIC                                   # XZArchive.swift:39
    yield "header magic", xz_stream(header=xz_header(magic=b"\xFD7zXZ\x01")), XZ_MAGIC   # XZStreamHeader.swift:35
    yield "header flags crc", xz_stream(header=xz_header(bad_crc=True)), XZ_INFO_CRC     # :43
    yield "header flags byte 0", xz_stream(header=xz_header(flags0=1)), XZ_FIELD         # :48
    yield "header check type high nibble", xz_stream(header=xz_header(check=0x11)), XZ_FIELD
    yield "header unknown check type 2", xz_stream(header=xz_header(check=2)), XZ_FIELD   # :52-56
    yield "block flags reserved bit", xz_stream(block_kw=dict(flags=0x04)), XZ_FIELD     # XZBlock.swift:28 (before the CRC test)
    yield "filter id 0x22", xz_stream(block_kw=dict(filt=b"\x22\x01\x10")), XZ_FILTER    # :59
    yield "lzma2 properties size 2", xz_stream(block_kw=dict(filt=b"\x21\x02\x10\x00")), LZMA2_DICT   # :46
    yield "delta properties size 2", xz_stream(block_kw=dict(flags=0x01, filt=b"\x03\x02\x00\x00\x21\x01\x10")), XZ_FIELD   # :54
    yield "header padding not zero", xz_stream(block_kw=dict(hdr_pad_byte=1)), XZ_PADDING   # :66 (before the CRC test)
    yield "block header crc", xz_stream(block_kw=dict(bad_hdr_crc=True)), XZ_INFO_CRC    # :72
    yield "compressed size field", xz_stream(block_kw=dict(comp=len(LZMA2) + 1)), XZ_DATA_SIZE   # :79
    yield "uncompressed size field", xz_stream(block_kw=dict(uncomp=len(RAW) - 1)), XZ_DATA_SIZE
    yield "vli with a zero continuation byte", xz_stream(block_kw=dict(comp=b"\x80\x00")), XZ_VLI   # LittleEndianByteReader+XZ.swift:20
    if len(LZMA2) % 4:
        yield "block padding not zero", xz_stream(block_kw=dict(block_pad_byte=7)), XZ_PADDING   # XZBlock.swift:89
    yield "check value", xz_stream(block_kw=dict(check_bytes=b"\0\0\0\1")), XZ_CHECK      # XZArchive.swift:112 (payload returned)
    yield "index record count", xz_stream(index_kw=dict(count=2)), XZ_FIELD              # :136
    blk, unp = xz_block()
    yield "index unpadded size", xz_stream(records=[(unp + 1, len(RAW))]), XZ_FIELD      # :141
    yield "index uncompressed size", xz_stream(records=[(unp, len(RAW) + 1)]), XZ_DATA_SIZE   # :145
    yield "index crc", xz_stream(index_kw=dict(bad_crc=True)), XZ_INFO_CRC               # :163
    yield "index padding not zero", xz_stream(index_kw=dict(pad_byte=5)), XZ_PADDING     # :154
    yield "footer crc", xz_stream(footer_kw=dict(bad_crc=True)), XZ_INFO_CRC             # :177
    yield "footer backward size", xz_stream(footer_kw=dict(backward=7)), XZ_FIELD        # :180
    yield "footer flags differ from the header", xz_stream(footer_kw=dict(flags=b"\x00\x04")), XZ_FIELD   # :184
    yield "footer flags reserved byte", xz_stream(footer_kw=dict(flags=b"\x01\x01")), XZ_FIELD
    yield "footer magic", xz_stream(footer_kw=dict(magic=b"YY")), XZ_MAGIC               # :190
    yield "two streams, 4 bytes of padding", GOOD + bytes(4) + GOOD, OK
    yield "two streams, 8 bytes of padding", GOOD + bytes(8) + GOOD, OK
    yield "two streams, 3 bytes of padding", GOOD + bytes(3) + GOOD, XZ_PADDING          # :203
    yield "4 trailing zero bytes", GOOD + bytes(4), OK                                   # :208-212 (paddingBytes % 4 == 3 at EOF)
    yield "5 trailing zero bytes", GOOD + bytes(5), XZ_PADDING
    yield "2 trailing zero bytes", GOOD + bytes(2), XZ_PADDING
    yield "garbage after the stream", GOOD + b"\x01", XZ_MAGIC                           # next loop turn: < 32 bytes left


def lz4_frame(flg=0x60, bd=0x40, blocks=(), content_size=None, dict_id=None, hc=None, end=b"\0\0\0\0", content_checksum=None,
              magic=b"\x04\x22\x4D\x18"):
    desc = bytes([flg, bd])
    if content_size is not None:
        desc += struct.pack("<Q", content_size)
    if dict_id is not None:
        desc += struct.pack("<I", dict_id)
    h = (xxh32(desc) >> 8) & 0xFF if hc is None else hc
    out = magic + desc + bytes([h])
    for b in blocks:
        out += b
    out += end
    if content_checksum is not None:
        out += struct.pack("<I", content_checksum)
    return out


def lz4_block(raw, stored=False, checksum=None, size_field=None):
    body = raw if stored else H.lz4_block_compress(raw)
    mark = (len(body) if size_field is None else size_field) | (0x80000000 if stored else 0)
    out = struct.pack("<I", mark) + body
    if checksum is not None:
        out += struct.pack("<I", xxh32(body) if checksum is True else checksum)
    return out


TXT = H.textlike(2000, 78)


def lz4_cases():
    yield "good", lz4_frame(blocks=[lz4_block(TXT)]), OK
    yield "good, stored block + checksums + size", lz4_frame(flg=0x7C, blocks=[lz4_block(TXT, stored=True, checksum=True)], content_size=len(TXT),
                                                             content_checksum=xxh32(TXT)), OK
    yield "empty frame", lz4_frame(), OK
    yield "magic only", b"\x04\x22\x4D\x18", TRUNC                                        # LZ4.swift:192 (< 7 bytes of frame)
    yield "version 2", lz4_frame(flg=0xA0), CORRUPT                                       # :198
    yield "reserved FLG bit", lz4_frame(flg=0x62), CORRUPT
    yield "BD reserved bits / unknown size", lz4_frame(bd=0x41), CORRUPT                  # :226
    yield "BD 0x30", lz4_frame(bd=0x30), CORRUPT
    yield "header checksum", lz4_frame(hc=0x00 if ((xxh32(bytes([0x60, 0x40])) >> 8) & 0xFF) else 0x01), CORRUPT   # :268
    yield "content size flagged, 12 bytes left", b"\x04\x22\x4D\x18" + bytes([0x68, 0x40]) + bytes(12), TRUNC   # :233
    yield "content size mismatch", lz4_frame(flg=0x68, blocks=[lz4_block(TXT)], content_size=len(TXT) + 1), CORRUPT   # :319
    yield "dict id present without a dictionary", lz4_frame(flg=0x61, dict_id=5), CORRUPT   # :249
    yield "block larger than the maximum", lz4_frame(blocks=[lz4_block(TXT, size_field=65537)]), CORRUPT   # :288
    yield "block checksum", lz4_frame(flg=0x70, blocks=[lz4_block(TXT, checksum=1)]), CORRUPT   # :297
    yield "missing end mark", lz4_frame(blocks=[lz4_block(TXT)], end=b""), TRUNC          # the block needs size + 4 more bytes (:291)
    yield "content checksum missing", lz4_frame(flg=0x64, blocks=[lz4_block(TXT)]), TRUNC   # :323
    yield "content checksum mismatch", lz4_frame(flg=0x64, blocks=[lz4_block(TXT)], content_checksum=1), MISMATCH   # :325
    yield "skippable frame, then a frame", b"\x50\x2A\x4D\x18" + struct.pack("<I", 3) + b"abc" + lz4_frame(blocks=[lz4_block(TXT)]), OK
    yield "skippable frame, size beyond the data", b"\x5F\x2A\x4D\x18" + struct.pack("<I", 9) + b"abc", TRUNC   # :152
    yield "legacy frame", b"\x02\x21\x4C\x18" + struct.pack("<I", len(H.lz4_block_compress(TXT))) + H.lz4_block_compress(TXT), OK
    yield "legacy frame, truncated block", b"\x02\x21\x4C\x18" + struct.pack("<I", 5000) + b"abc", TRUNC   # :177
    yield "unknown magic", b"\x05\x22\x4D\x18" + bytes(20), CORRUPT


XZ = list(xz_cases())
LZ = list(lz4_cases())


@pytest.mark.parametrize("name,blob,status", XZ, ids=[c[0] for c in XZ])
def test_xz_framing_oracle(oracle, name, blob, status):
    st, out, _ = oracle.xz_unarchive(blob)
    assert st == status, (name, st)
    if status == OK:
        assert out == RAW * (2 if name.startswith("two streams") else 1)
    if status == XZ_CHECK:
        assert out == RAW


@pytest.mark.parametrize("name,blob,status", LZ, ids=[c[0] for c in LZ])
def test_lz4_framing_oracle(oracle, name, blob, status):
    st, out, _ = oracle.lz4_decompress(blob)
    assert st == status, (name, st)
    if status in (OK, MISMATCH) and name not in ("empty frame",):
        assert out == TXT


@pytest.mark.gpu
def test_xz_and_lz4_framing_product():
    import swcompression_b200 as S
    for cases, fn, want_ok in ((XZ, S.XZArchive.unarchive, None), (LZ, S.LZ4.decompress, None)):
        for name, blob, status in cases:
            try:
                out = fn(blob)
                got = 0
            except S.SWCompressionError as e:
                got, out = e.code, e.payload
            assert got == status, (name, got, status)
            if status == OK and name != "empty frame":
                assert out == (RAW * (2 if name.startswith("two streams") else 1) if cases is XZ else TXT), name

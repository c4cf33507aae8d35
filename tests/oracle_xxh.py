"""Pure-Python xxHash32 (seed 0) used to *build* LZ4 frames in tests, independent of both the oracle and the product."""


def _rotl(v, s):
    return ((v << s) | (v >> (32 - s))) & 0xFFFFFFFF


P1, P2, P3, P4, P5 = 0x9E3779B1, 0x85EBCA77, 0xC2B2AE3D, 0x27D4EB2F, 0x165667B1


def xxh32(data):
    n = len(data)
    i = 0
    M = 0xFFFFFFFF
    if n >= 16:
        a = [(P1 + P2) & M, P2, 0, (-P1) & M]
        while n - i >= 16:
            for j in range(4):
                lane = int.from_bytes(data[i + 4 * j:i + 4 * j + 4], "little")
                a[j] = (_rotl((a[j] + lane * P2) & M, 13) * P1) & M
            i += 16
        acc = (_rotl(a[0], 1) + _rotl(a[1], 7) + _rotl(a[2], 12) + _rotl(a[3], 18)) & M
    else:
        acc = P5
    acc = (acc + n) & M
    while n - i >= 4:
        lane = int.from_bytes(data[i:i + 4], "little")
        acc = (_rotl((acc + lane * P3) & M, 17) * P4) & M
        i += 4
    while n - i >= 1:
        acc = (_rotl((acc + data[i] * P5) & M, 11) * P1) & M
        i += 1
    acc ^= acc >> 15
    acc = (acc * P2) & M
    acc ^= acc >> 13
    acc = (acc * P3) & M
    acc ^= acc >> 16
    return acc

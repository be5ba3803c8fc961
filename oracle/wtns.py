"""`.wtns` writer / reader of the oracle (snarkjs 0.3.59 `wtns_utils.writeBin`
layout [un-vendored]): "wtns" | u32 version=2 | u32
nSections=2 | section 1 (id=1, size=40): u32 n8=32, prime LE, u32 nWitness |
section 2 (id=2, size=32*nWitness): values, 32-byte little-endian canonical.
Test infrastructure only."""
import struct

from .field import P


def wtns_bytes(witness):
    n = len(witness)
    out = bytearray()
    out += b"wtns"
    out += struct.pack("<II", 2, 2)
    out += struct.pack("<IQ", 1, 40)
    out += struct.pack("<I", 32) + P.to_bytes(32, "little") + struct.pack("<I", n)
    out += struct.pack("<IQ", 2, 32 * n)
    for v in witness:
        out += int(v).to_bytes(32, "little")
    return bytes(out)


def parse_wtns(b):
    assert b[:4] == b"wtns"
    ver, nsec = struct.unpack_from("<II", b, 4)
    assert (ver, nsec) == (2, 2)
    sid, size = struct.unpack_from("<IQ", b, 12)
    assert (sid, size) == (1, 40)
    n8, = struct.unpack_from("<I", b, 24)
    prime = int.from_bytes(b[28:60], "little")
    n, = struct.unpack_from("<I", b, 60)
    sid, size = struct.unpack_from("<IQ", b, 64)
    assert sid == 2 and size == 32 * n and n8 == 32 and prime == P
    return [int.from_bytes(b[76 + 32 * i: 108 + 32 * i], "little") for i in range(n)]

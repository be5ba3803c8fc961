"""The ladder kernels of the signature pipeline (csrc/k_ladder.cu) against big-integer arithmetic, point by point.

k_ed_ladder walks the three chains of a signature (doublings of 8A, the accumulators of the two EscalarMulAny
segments, the fixed-base accumulator) on twisted-Edwards extended coordinates when 8A is a curve point and on
projective triples of the templates' own rational maps when it is not (the key of an L1 slot,
/root/reference/src/rollup-tx.circom:447-476).  The points it leaves (hzw_debug_read, which = 1) must be the
affine Montgomery points MontgomeryDouble / MontgomeryAdd of circomlib produce; the mixed batch holds both kinds."""
import ctypes

import pytest

from conftest import MIXED32

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
A = 168698
S = 0x0e0ff35005d6e54b2c7ac0ce84e6117774a5f6ca321ea1176c7be2c1d8b2d768      # sqrt(-168700)
RI = pow(1 << 256, -1, P)
B8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
      16950150798460657717958625567821834550301663161624707787222815936182638968203)
ARR = 704
REC = 4 * ARR + 96 + 176
MID = REC + 4


def inv(x):
    return pow(x % P, -1, P)


def m_dbl(p):
    u, v = p
    lam = (3 * u * u + 2 * A * u + 1) * inv(2 * v) % P
    u3 = (lam * lam - A - 2 * u) % P
    return u3, (lam * (u - u3) - v) % P


def m_add(p, q):
    (u1, v1), (u2, v2) = p, q
    lam = (v2 - v1) * inv(u2 - u1) % P
    u3 = (lam * lam - A - u1 - u2) % P
    return u3, (lam * (u1 - u3) - v1) % P


def affine(pt, on_curve):
    """(u, v) of a stored triple: (X', Y, Z) of the twisted form or (U, V, W), both in the Montgomery domain"""
    x, y, z = (c * RI % P for c in pt)
    if on_curve:
        return (z + y) * inv(z - y) % P, S * (z + y) * z % P * inv((z - y) * x) % P
    return x * inv(z) % P, y * inv(z) % P


@pytest.mark.gpu
def test_ladder_points_equal_the_montgomery_chain(built, mixed32):
    from circuits_b200.witness_calculator import WitnessCalculator, _addr
    assert (S * S + 168700) % P == 0
    wc = WitnessCalculator(*MIXED32)
    wc.calculateResident(wc.packInput(mixed32.getInput()), True)
    lib = wc._lib
    lib.hzw_debug_read.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]

    def rd(which, g, first, count):
        buf = bytearray(32 * count)
        assert lib.hzw_debug_read(wc._ctx, which, g, first, count, _addr(buf)) == 0
        return [int.from_bytes(buf[32 * i:32 * i + 32], "little") for i in range(count)]

    kinds = set()
    for g in range(MIXED32[0]):
        words = []
        for v in rd(0, g, REC, 4):
            words += [(v >> (32 * k)) & 0xffffffff for k in range(8)]
        ok, e, form = words[0], words[4:12], words[22]
        px, py = rd(0, g, MID, 2)
        if px == 0:
            px, py = B8                                      # EscalarMulAny's zero-point substitution
        on_curve = (168700 * px * px + py * py - 1 - 168696 * px * px * py * py) % P == 0
        kinds.add(on_curve)
        assert ok == 1, "signature %d fell back to the templates' divisions" % g
        # a block of 8 signatures walks one form: the twisted one only if all of them are curve points
        assert form in (0, 1) and (form == 0 or on_curve)
        pts = rd(1, g, 0, 1776)
        u = (1 + py) * inv(1 - py) % P
        D = [(u, u * inv(px) % P)]
        for t in range(253):
            D.append(m_dbl(D[t]))
        for t in range(254):
            assert affine(pts[3 * t:3 * t + 3], form == 1) == D[t], ("D", g, t)
        acc = D[0]
        for t in range(1, 254):
            at = 762 + 3 * (t - 1)
            assert affine(pts[at:at + 3], form == 1) == acc, ("A", g, t)
            if t == 148:
                acc = D[148]
                continue
            r = m_add(D[t], acc)
            if (e[t >> 5] >> (t & 31)) & 1:
                acc = r
        assert affine(pts[1770:1773], form == 1) == acc
    assert kinds == {True, False}, "the batch should hold signatures on and off the curve"

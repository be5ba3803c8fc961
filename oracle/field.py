"""BN254 scalar field Fr helpers (oracle; test infrastructure only).

p is the prime the reference builds its field for
(/root/reference/tools/helpers/actions.js:209).  Semantics follow the ffwasm /
ffiasm field libraries the circom-generated calculators run on [un-vendored]:
values are canonical residues, `x / 0` evaluates to 0, and ordered comparisons
use circom's signed interpretation (x > (p-1)/2 is negative).
"""

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
HALF = (P - 1) >> 1


def inv(x):
    """Field inverse with the calculators' convention inv(0) = 0."""
    x %= P
    if x == 0:
        return 0
    return pow(x, -1, P)


def div(a, b):
    return (a % P) * inv(b) % P


def is_neg(x):
    """circom signed compare: x < 0  <=>  canonical x > (p-1)/2."""
    return (x % P) > HALF


def to_le32(x):
    return int(x % P).to_bytes(32, "little")


def from_le32(b):
    return int.from_bytes(b, "little")

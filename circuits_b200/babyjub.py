"""BabyJubJub + EdDSA-Poseidon signing for the host-side batch builder.

Host mirror of what circomlib's `babyjub.js` / `eddsa.js` give the reference's
input generation (HermezAccount.signTx in @hermeznetwork/commonjs [un-vendored];
used at /root/reference/tools/helpers/gen-inputs-utils.js:60).  Twisted Edwards
a*x^2 + y^2 = 1 + d*x^2*y^2, a = 168700, d = 168696.
"""
from .field import P, inv
from .poseidon import poseidon

A = 168700
D = 168696
BASE8 = (5299619240641551281634865583518297030282874472190772894086521144482721001553,
         16950150798460657717958625567821834550301663161624707787222815936182638968203)
SUBORDER = 2736030358979909402780800718157159386076813972158567259200215660948447373041


def add(p1, p2):
    x1, y1 = p1
    x2, y2 = p2
    beta = x1 * y2 % P
    gamma = y1 * x2 % P
    delta = (-A * x1 + y1) * (x2 + y2) % P
    tau = beta * gamma % P
    x3 = (beta + gamma) * inv(1 + D * tau) % P
    y3 = (delta + A * beta - gamma) * inv(1 - D * tau) % P
    return (x3, y3)


def mul(k, pt):
    acc = (0, 1)
    while k:
        if k & 1:
            acc = add(acc, pt)
        pt = add(pt, pt)
        k >>= 1
    return acc


def on_curve(pt):
    x, y = pt
    x2, y2 = x * x % P, y * y % P
    return (A * x2 + y2 - 1 - D * x2 * y2) % P == 0


def pack_bits(pt):
    """256 compressed bits, LSB first: ay in bits 0..253 (bit 254 = 0),
    sign (x > (p-1)/2) in bit 255 -- the `fromBjjCompressed` input layout
    (/root/reference/src/lib/utils-bjj.circom:12-28)."""
    x, y = pt
    sign = 1 if x > (P - 1) // 2 else 0
    bits = [(y >> i) & 1 for i in range(255)] + [sign]
    return bits


def sign_of(x):
    return 1 if x > (P - 1) // 2 else 0


class KeyPair:
    """EdDSA key: secret scalar s = 8*a (a pruned private scalar), A = a*B8."""

    def __init__(self, a):
        self.a = a % SUBORDER or 1
        self.s = self.a * 8
        self.pub = mul(self.a, BASE8)

    def sign(self, msg, nonce_seed):
        r = (nonce_seed % (SUBORDER - 1)) + 1
        R8 = mul(r, BASE8)
        hm = poseidon([R8[0], R8[1], self.pub[0], self.pub[1], msg])
        S = (r + hm * self.s) % SUBORDER
        return R8, S


def verify(pub, msg, R8, S):
    hm = poseidon([R8[0], R8[1], pub[0], pub[1], msg])
    a8 = mul(8, pub)
    return mul(S, BASE8) == add(R8, mul(hm, a8))

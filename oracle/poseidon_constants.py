"""Poseidon round constants and MDS matrices for circomlib 0.5.2 (oracle).

circomlib's `poseidon_constants.circom` [un-vendored, pinned version 0.5.2 in
/root/reference/package-lock.json] is a dump of the output of the Hades
reference script `generate_parameters_grain.sage 1 0 254 t 8 R_P p`.  The
script's published algorithm is restated here (Grain LFSR in self-shrinking
mode); tests/test_oracle_kat.py pins the result against the upstream
known-answer vectors for t = 2, 3, 5, 6, 7.

Used by: /root/reference/src/decode-tx.circom:275 (Poseidon(6)),
/root/reference/src/lib/hash-state.circom:32 (Poseidon(4)), and through
circomlib's SMT (Poseidon(2), Poseidon(3)) and EdDSA (Poseidon(5)) templates.
"""
from functools import lru_cache

from .field import P

N_ROUNDS_F = 8
N_ROUNDS_P = [56, 57, 56, 60, 60, 63, 64, 63]  # indexed by t-2


def _grain_stream(n, t, rf, rp):
    state = []

    def push(v, w):
        state.extend((v >> (w - 1 - i)) & 1 for i in range(w))

    push(1, 2)      # field type: prime field
    push(0, 4)      # s-box: x^alpha
    push(n, 12)
    push(t, 12)
    push(rf, 10)
    push(rp, 10)
    state.extend([1] * 30)
    assert len(state) == 80

    def step():
        b = state[62] ^ state[51] ^ state[38] ^ state[23] ^ state[13] ^ state[0]
        state.pop(0)
        state.append(b)
        return b

    for _ in range(160):
        step()
    while True:
        b = step()
        while b == 0:
            step()
            b = step()
        yield step()


@lru_cache(maxsize=None)
def poseidon_params(t):
    """Return (C, M, nRoundsF, nRoundsP): C flat list of t*(RF+RP), M t x t."""
    n = 254
    rf, rp = N_ROUNDS_F, N_ROUNDS_P[t - 2]
    g = _grain_stream(n, t, rf, rp)

    def draw():
        v = 0
        for _ in range(n):
            v = (v << 1) | next(g)
        return v

    C = []
    for _ in range((rf + rp) * t):
        r = draw()
        while r >= P:
            r = draw()
        C.append(r)
    while True:
        rl = [draw() % P for _ in range(2 * t)]
        while len(set(rl)) != len(rl):
            rl = [draw() % P for _ in range(2 * t)]
        xs, ys = rl[:t], rl[t:]
        if all((x + y) % P for x in xs for y in ys):
            break
    M = [[pow((xs[i] + ys[j]) % P, -1, P) for j in range(t)] for i in range(t)]
    return C, M, rf, rp


def poseidon_hash(inputs):
    """Plain (value-only) Poseidon of circomlib 0.5.2: out = state[0]."""
    t = len(inputs) + 1
    C, M, rf, rp = poseidon_params(t)
    s = [0] + [x % P for x in inputs]
    for r in range(rf + rp):
        s = [(a + C[r * t + i]) % P for i, a in enumerate(s)]
        if r < rf // 2 or r >= rf // 2 + rp:
            s = [pow(a, 5, P) for a in s]
        else:
            s[0] = pow(s[0], 5, P)
        s = [sum(M[i][j] * s[j] for j in range(t)) % P for i in range(t)]
    return s[0]

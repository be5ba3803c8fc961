"""BN254 Fr constants shared by the host-side modules (product code)."""

P = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R_MONT = (1 << 256) % P


def inv(x):
    x %= P
    return 0 if x == 0 else pow(x, -1, P)

"""Test-side writers of the circom toolchain's files (the product only READS them: circuits_b200/csrc/formats.hpp).

`write_sym`   a circom 0.5.x `.sym` (`labelIdx,varIdx,componentIdx,name`) for a made-up compiler layout: the signals
              of the circuit renumbered by a seeded permutation, a share of them eliminated (`varIdx = -1`, what `-r`
              does), names that share an entry here sharing one there.
`write_r1cs`  an iden3 `.r1cs` (v1) from a list of constraints ((A, B, C) with each side a dict wire -> coefficient).
`gadget_constraints`  the constraints of every Num2Bits / IsZero component found by name in a symbol table
              (circomlib bitify.circom: out[k] * (out[k] - 1) = 0, sum 2^k out[k] = in; comparators.circom:
              in * inv = 1 - out, in * out = 0), kept only where the reference witness satisfies them."""
import random
import re
import struct

from oracle.field import P


def write_sym(path, symbols, n_inputs, seed=7, drop=0.3):
    """symbols: name -> own index.  Returns (var_of_own: own index -> compiler index or None, n_vars)."""
    rng = random.Random(seed)
    own = sorted(set(symbols.values()))
    keep = [i for i in own if i <= 1 + n_inputs or rng.random() >= drop]      # outputs and inputs are never eliminated
    head = [i for i in keep if i <= 1 + n_inputs]
    tail = [i for i in keep if i > 1 + n_inputs]
    rng.shuffle(tail)
    var_of_own = {i: v + 1 for v, i in enumerate(head + tail)}
    with open(path, "w") as f:
        for label, (name, idx) in enumerate(sorted(symbols.items(), key=lambda kv: (kv[1], kv[0])), 1):
            f.write("%d,%d,%d,%s\n" % (label, var_of_own.get(idx, -1), 0, name))
    return var_of_own, len(keep) + 1


def write_r1cs(path, constraints, n_wires, n_pub_out=1, n_pub_in=0, n_prv_in=0):
    def lc(d):
        out = struct.pack("<I", len(d))
        for wire in sorted(d):
            out += struct.pack("<I", wire) + (d[wire] % P).to_bytes(32, "little")
        return out
    header = struct.pack("<I", 32) + P.to_bytes(32, "little") + struct.pack("<IIIIQI", n_wires, n_pub_out, n_pub_in, n_prv_in,
                                                                              n_wires, len(constraints))
    body = b"".join(lc(a) + lc(b) + lc(c) for a, b, c in constraints)
    w2l = b"".join(struct.pack("<Q", i) for i in range(n_wires))
    with open(path, "wb") as f:
        f.write(b"r1cs" + struct.pack("<II", 1, 3))
        # sections deliberately out of order (constraints first): readers must not assume the header comes first
        for typ, data in ((2, body), (1, header), (3, w2l)):
            f.write(struct.pack("<IQ", typ, len(data)) + data)


def gadget_constraints(symbols, witness, index_of=lambda i: i):
    """Constraints over wires `index_of(own index)`; `witness` is indexed by own index (used to keep true ones only)."""
    comps = {}
    for name, idx in symbols.items():
        pre, _, leaf = name.rpartition(".")
        comps.setdefault(pre, {})[leaf] = idx
    cons = []
    for pre, sigs in comps.items():
        outs = sorted((int(m.group(1)), i) for k, i in sigs.items() for m in [re.fullmatch(r"out\[(\d+)\]", k)] if m)
        if "in" in sigs and outs and len(sigs) == 1 + len(outs) and [k for k, _ in outs] == list(range(len(outs))):
            if sum(witness[i] << k for k, i in outs) % P != witness[sigs["in"]] or any(witness[i] > 1 for _, i in outs):
                continue
            if any(index_of(i) is None for _, i in outs) or index_of(sigs["in"]) is None:
                continue
            for _, i in outs:
                cons.append(({index_of(i): 1}, {index_of(i): 1, 0: -1}, {}))
            s = {}
            for k, i in outs:
                s[index_of(i)] = (s.get(index_of(i), 0) + (1 << k)) % P
            cons.append(({}, {}, dict_sub(s, {index_of(sigs["in"]): 1})))
        elif set(sigs) == {"in", "out", "inv"}:
            a, o, v = witness[sigs["in"]], witness[sigs["out"]], witness[sigs["inv"]]
            if (a * v - (1 - o)) % P or (a * o) % P or any(index_of(sigs[k]) is None for k in ("in", "out", "inv")):
                continue
            cons.append(({index_of(sigs["in"]): 1}, {index_of(sigs["inv"]): 1}, dict_sub({0: 1}, {index_of(sigs["out"]): 1})))
            cons.append(({index_of(sigs["in"]): 1}, {index_of(sigs["out"]): 1}, {}))
    return cons


def dict_sub(a, b):
    out = dict(a)
    for k, v in b.items():
        out[k] = (out.get(k, 0) - v) % P
    return {k: v for k, v in out.items() if v}

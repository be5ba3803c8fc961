#!/usr/bin/env python3
"""Extract the literal golden vectors of the reference's own mocha tests into
tests/golden/reference_vectors.json.

Run in the development container only (reads /root/reference, which does not
exist on the GPU box):   python tests/golden/extract_reference_vectors.py

Sources (all literal in/out tables):
  test/lib/decode-float.test.js:28-38       float40 vectors
  test/rollup-tx-states.test.js:38-625      RollupTxStates tables (mini JS interpreter below)
  test/fee-accumulator.test.js:28-113       FeeAccumulator(16) vectors (all 7)
  test/balance-updater.test.js:31-190       BalanceUpdater cases (fees via the fee table, which
                                            is itself pinned by test/rollup-main.test.js:555,616)
The interpreter understands the handful of statement shapes those files use:
`const|let input = {..}`, `input.k = v`, `output = {..}`, `output.k = v`,
`calculateWitness(input` (snapshot), `assertOut(w, output)` (emit a case) and
a `try { calculateWitness } catch` expecting "Constraint doesn't match".
"""
import json
import os
import re
import sys

REF = "/root/reference/test"
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from circuits_b200.feetable import compute_fee  # noqa: E402  (host-side fee table, pinned separately)


class AttrDict(dict):
    __getattr__ = dict.__getitem__


def js_eval(expr, env):
    e = expr.strip().rstrip(",").strip()
    e = re.sub(r"Scalar\.(\w+)", r"Scalar_\1", e)
    e = re.sub(r"Constants\.(\w+)", r"Constants_\1", e)
    m = re.match(r"^(.*?)\?(.*?):(.*)$", e)
    if m:
        return js_eval(m.group(2), env) if js_eval(m.group(1), env) else js_eval(m.group(3), env)
    e = e.replace("===", "==").replace("!==", "!=")
    return eval(e, {"__builtins__": {}}, env)


BASE_ENV = dict(
    Scalar_e=lambda x: int(x), Scalar_add=lambda a, b: int(a) + int(b),
    Scalar_sub=lambda a, b: int(a) - int(b), Scalar_mul=lambda a, b: int(a) * int(b),
    Scalar_isZero=lambda a: int(a) == 0, Scalar_fromString=lambda s, r=10: int(s, r),
    Scalar_pow=lambda a, b: int(a) ** int(b), Scalar_shl=lambda a, b: int(a) << int(b),
    Scalar_gt=lambda a, b: int(a) > int(b),
    Constants_exitIdx=1, Constants_nullIdx=0, Constants_nullEthAddr=(1 << 160) - 1,
    computeFee=lambda a, s: compute_fee(int(a), int(s)), true=True, false=False,
)


def parse_object(lines, i, env):
    """Parse `{ k: v, ... }` starting at lines[i] (which contains the `{`)."""
    obj = AttrDict()
    i += 1
    while not lines[i].strip().startswith("}"):
        ln = lines[i].split("//")[0].strip()
        if ln:
            k, v = ln.split(":", 1)
            obj[k.strip()] = js_eval(v, env)
        i += 1
    return obj, i


def interpret_it_blocks(path):
    src = open(path).read().split("\n")
    cases = []
    i = 0
    title = None
    env = dict(BASE_ENV)
    snap = None
    in_try = False
    while i < len(src):
        ln = src[i].strip()
        m = re.match(r'it\("(.*?)"', ln)
        if m:
            title = m.group(1)
            env = dict(BASE_ENV)
            snap = None
            in_try = False
        elif re.match(r"(const|let)?\s*(input|output)\s*=\s*\{", ln):
            name = re.match(r"(?:const|let)?\s*(\w+)\s*=", ln).group(1)
            env[name], i = parse_object(src, i, env)
        elif re.match(r"(const|let)\s+(\w+)\s*=\s*[^{]+;$", ln) and "await" not in ln and "require" not in ln:
            m = re.match(r"(?:const|let)\s+(\w+)\s*=\s*(.*);$", ln)
            try:
                env[m.group(1)] = js_eval(m.group(2), env)
            except Exception:
                pass
        elif re.match(r"(input|output)\.(\w+)\s*=\s*(.*);$", ln):
            m = re.match(r"(input|output)\.(\w+)\s*=\s*(.*);$", ln)
            env[m.group(1)][m.group(2)] = js_eval(m.group(3), env)
        elif ln.startswith("try"):
            in_try = True
        elif "calculateWitness(input" in ln:
            snap = dict(env["input"])
            if in_try:
                cases.append(dict(title=title, input=snap, error="Constraint doesn't match"))
                in_try = False
        elif "assertOut(w, output)" in ln:
            cases.append(dict(title=title, input=snap, output=dict(env["output"])))
        i += 1
    return cases


def fee_accumulator_vectors(path):
    src = open(path).read()
    body = src[src.index("const testVectors = ["):]
    body = body[body.index("["): body.index("];") + 1]
    body = re.sub(r"//.*", "", body)
    body = re.sub(r"(\w+):", r'"\1":', body)
    body = re.sub(r",(\s*[}\]])", r"\1", body)
    vecs = json.loads(body)
    out = []
    for v in vecs:
        exp = (v.get("out") or {}).get("accFeeOut") or (v.get("out") or {}).get("accFeeIn")
        out.append(dict(input=v["input"], accFeeOut=exp if exp is not None else v["input"]["accFeeIn"]))
    return out


def decode_float_vectors(path):
    src = open(path).read()
    out = []
    for m in re.finditer(r"\[\s*([0-9x* A-Fa-f+]+?),\s*\"(\d+)\"\s*\]", src):
        out.append([eval(m.group(1), {"__builtins__": {}}), int(m.group(2))])
    return out


def main():
    res = dict(
        decode_float=decode_float_vectors(os.path.join(REF, "lib/decode-float.test.js")),
        rollup_tx_states=interpret_it_blocks(os.path.join(REF, "rollup-tx-states.test.js")),
        fee_accumulator=fee_accumulator_vectors(os.path.join(REF, "fee-accumulator.test.js")),
        balance_updater=interpret_it_blocks(os.path.join(REF, "balance-updater.test.js")),
    )

    def stringify(o):
        if isinstance(o, bool):
            return int(o)
        if isinstance(o, int):
            return str(o)
        if isinstance(o, dict):
            return {k: stringify(v) for k, v in o.items()}
        if isinstance(o, list):
            return [stringify(v) for v in o]
        return o
    dst = os.path.join(os.path.dirname(__file__), "reference_vectors.json")
    with open(dst, "w") as f:
        json.dump(stringify(res), f, indent=1, sort_keys=True)
    print({k: len(v) for k, v in res.items()}, "->", dst)


if __name__ == "__main__":
    main()

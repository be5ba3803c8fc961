"""CPU oracle for the Hermez rollup-main / withdraw witness path.

TEST INFRASTRUCTURE ONLY.  This package is a big-integer restatement of the
reference's circuit semantics (hermeznetwork/circuits `src/*.circom` plus the
circomlib 0.5.2 templates those files include).  Only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of
`bench.py` may import it; the product path (`circuits_b200/`) never does and
fails loudly when its CUDA library is missing.

Parity status
-------------
* In-repo templates (`/root/reference/src/**/*.circom`) are restated from the
  source files, each function citing file:line.
* circomlib 0.5.2, circom 0.5.45 and circom_runtime 0.1.12 are npm
  dependencies that are NOT vendored under /root/reference (package.json:24-31,
  no node_modules) and cannot be installed (no network, no node).  Their
  templates are restated from the published algorithms and pinned by
  known-answer vectors: Poseidon (circomlib test/poseidonjs.js and
  go-iden3-crypto vectors), BabyJubJub group identities, SHA-256 against
  hashlib, the Tonelli-Shanks constants of pointbits.circom, plus every
  literal golden vector the reference's own tests hold.
* Witness VALUES are pinned as above.  Witness LAYOUT (signal index <-> name)
  is decided by the circom compiler, which cannot run here: "layout parity
  unpinned" -- the canonical order defined in `oracle/circuit.py` is used by
  both the oracle and the CUDA path, and a circom `.sym` map can be applied on
  top by the host layer.
"""

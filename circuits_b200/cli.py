"""Command line of the witness generator: the place of the circom-generated binary in the reference's tooling.

    python -m circuits_b200.cli --nTx 2048 --nLevels 32 --maxL1Tx 256 --maxFeeTx 64 input.json witness.wtns

mirrors `./circuit input.json witness.wtns` (/root/reference/tools/helpers/actions.js:139-143: the reference runs the
calculator circom compiled for the circuit; here the circuit is named by its four parameters) and, with a `.json`
output name, `snarkjs wtns export json`'s array of decimal strings.  The input file is what
`tools/generate-input.js` writes: one JSON object keyed by the main component's input signals
(/root/reference/src/rollup-main.circom:105-196), values as numbers, decimal strings or nested arrays.
`--circuit withdraw` takes the inputs of /root/reference/src/withdraw.circom (only --nLevels matters).
Exit status: 0, or 1 with the reference's message on stderr (`--sanity` turns the `===` checks on, as a binary built
with SANITY_CHECK does).
"""
import argparse
import json
import sys


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m circuits_b200.cli", description=__doc__.split("\n\n")[0])
    ap.add_argument("--circuit", default="rollup-main", choices=["rollup-main", "withdraw", "rollup-tx"])
    ap.add_argument("--sym", default=None, help="circom .sym of the circuit: the witness is written in the compiler's layout")
    ap.add_argument("--nTx", type=int, default=1)
    ap.add_argument("--nLevels", type=int, required=True)
    ap.add_argument("--maxL1Tx", type=int, default=0)
    ap.add_argument("--maxFeeTx", type=int, default=1)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--sanity", action="store_true", help="check every `===` of the circuit (sanityCheck)")
    ap.add_argument("--lib", default=None, help=argparse.SUPPRESS)      # tests only: the checker's library behind the same C ABI
    ap.add_argument("input", help="input.json")
    ap.add_argument("output", help="witness.wtns, or witness.json for the decimal-string array")
    args = ap.parse_args(argv)
    from .witness_calculator import WitnessCalculator, WitnessError
    with open(args.input) as f:
        inp = json.load(f)
    if not isinstance(inp, dict):
        sys.stderr.write("input must be a JSON object keyed by signal name\n")
        return 1
    try:
        wc = WitnessCalculator(args.nTx, args.nLevels, args.maxL1Tx, args.maxFeeTx, circuit=args.circuit,
                               device=args.device, lib_path=args.lib, sym=args.sym)
        if args.circuit == "rollup-main" and not args.output.endswith(".json"):
            # straight from device memory into the file, chunk by chunk: no copy of the witness in this process
            wc.calculateToWtns(inp, args.output, args.sanity)
            return 0
        w = wc.calculateBinWitness(inp, args.sanity)
    except WitnessError as e:
        sys.stderr.write("%s\n" % e)
        return 1
    if args.output.endswith(".json"):
        wb = bytes(w)
        with open(args.output, "w") as f:
            json.dump([str(int.from_bytes(wb[32 * i:32 * i + 32], "little")) for i in range(len(wb) // 32)], f)
    else:
        wc.writeWtns(w, args.output)
    return 0


if __name__ == "__main__":
    sys.exit(main())

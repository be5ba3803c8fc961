// circom_runtime 0.1.12 `WitnessCalculator` surface over libhzwitness.so (N-API addon hzwitness_napi.cc -> C ABI of
// include/hzwitness.h -> hand-written CUDA for sm_100a).  What the reference obtains from
//   const circuit = await tester(path, {reduceConstraints: false})      (/root/reference/test/rollup-main.test.js:52)
//   const w = await circuit.calculateWitness(input, {logOutput: false})  (/root/reference/test/helpers/helpers.js:149)
//   await circuit.assertOut(w, output); circuit.getSignal(w, name)        (helpers.js:151-154, 170)
// and from `./circuit input.json witness.wtns` (/root/reference/tools/helpers/actions.js:139-143).
//
//   const builder = require("hzwitness");
//   const wc = await builder({circuit: "rollup-main", nTx: 2048, nLevels: 32, maxL1Tx: 256, maxFeeTx: 64, nGpus: 8});
//   const bin = await wc.calculateBinWitness(bb.getInput(), true);   // Uint8Array(NVars * 32), 32-byte LE canonical
//   await wc.writeWtns(bin, "witness.wtns");
"use strict";
const path = require("path");
const native = require(path.join(__dirname, "build", "Release", "hzwitness_napi.node"));

const PRIME = 21888242871839275222246405745257275088548364400416034343698204186575808495617n;
const KINDS = { "rollup-main": 0, "withdraw": 1, "rollup-tx": 2 };

function flatten(v, out) {
    if (Array.isArray(v)) { for (const x of v) flatten(x, out); return; }
    out.push(((BigInt(v) % PRIME) + PRIME) % PRIME);
}
function toLE32(v, buf, at) {
    for (let b = 0; b < 32; b++) { buf[at + b] = Number(v & 0xffn); v >>= 8n; }
}
function fromLE32(buf, at) {
    let v = 0n;
    for (let b = 31; b >= 0; b--) v = (v << 8n) | BigInt(buf[at + b]);
    return v;
}

class WitnessCalculator {
    // options: {circuit, nTx, nLevels, maxL1Tx, maxFeeTx, device = 0, nGpus = 1, sym = null}
    constructor(options) {
        const o = Object.assign({ circuit: "rollup-main", nTx: 1, maxL1Tx: 0, maxFeeTx: 1, device: 0, nGpus: 1, sym: null }, options);
        if (!(o.circuit in KINDS)) throw new Error(`unknown circuit ${o.circuit}`);
        this.ctx = native.create(KINDS[o.circuit], o.nTx, o.nLevels, o.maxL1Tx, o.maxFeeTx, o.device, o.nGpus, o.sym);
        this.NVars = native.nvars(this.ctx);
        this.NInputs = native.ninputs(this.ctx);
        this.prime = PRIME;
        this.n64 = 4;
        this.n32 = 8;
    }

    // what circom_runtime's `_doCalculateWitness` does with `getSignalOffset32` + `setSignal`: keys are main input names
    packInput(input) {
        const buf = new Uint8Array(32 * this.NInputs);
        let seen = 0;
        for (const k of Object.keys(input)) {
            const [off, len] = native.inputIndex(this.ctx, k);          // throws "Signal k is not an input of the circuit."
            const vals = [];
            flatten(input[k], vals);
            if (vals.length !== len) throw new Error(`Invalid number of values for input signal ${k}: ${vals.length} != ${len}`);
            for (let i = 0; i < vals.length; i++) toLE32(vals[i], buf, 32 * (off + i));
            seen += len;
        }
        if (seen !== this.NInputs) throw new Error(`Not all inputs have been set. Only ${seen} out of ${this.NInputs}`);
        return buf;
    }

    async calculateBinWitness(input, sanityCheck) {
        const packed = input instanceof Uint8Array ? input : this.packInput(input);
        return native.calculate(this.ctx, packed, !!sanityCheck);         // rejects with "Constraint doesn't match <lhs> != <rhs> <where>"
    }

    async calculateWitness(input, sanityCheck) {
        const bin = await this.calculateBinWitness(input, sanityCheck);
        const w = new Array(this.NVars);
        for (let i = 0; i < this.NVars; i++) w[i] = fromLE32(bin, 32 * i);
        return w;
    }

    _bin(witness) {
        if (witness instanceof Uint8Array) return witness;
        const buf = new Uint8Array(32 * this.NVars);
        for (let i = 0; i < this.NVars; i++) toLE32(BigInt(witness[i]), buf, 32 * i);
        return buf;
    }

    // ---- circom tester helpers ------------------------------------------------------------------------------------
    async getSignal(witness, name) {
        return fromLE32(native.getSignal(this.ctx, this._bin(witness), name.startsWith("main.") ? name : "main." + name), 0);
    }

    async assertOut(witness, expected) {
        const bin = this._bin(witness);
        const check = async (name, want) => {
            if (Array.isArray(want)) { for (let i = 0; i < want.length; i++) await check(`${name}[${i}]`, want[i]); return; }
            const got = fromLE32(native.getSignal(this.ctx, bin, name), 0);
            const w = ((BigInt(want) % PRIME) + PRIME) % PRIME;
            if (got !== w) throw new Error(`${name}: ${got} != ${w}`);
        };
        for (const k of Object.keys(expected)) await check("main." + k, expected[k]);
    }

    async checkConstraints(witness, r1csPath) {
        const [n, bad, first] = native.checkConstraints(this.ctx, r1csPath, witness ? this._bin(witness) : null);
        if (bad) throw new Error(`Constraint doesn't match: ${bad} of ${n} constraints violated, first #${first}`);
    }

    async writeWtns(witness, file) { native.writeWtns(this.ctx, this._bin(witness), file); }
}

// circom_runtime exports `builder(code, options)`; here the "code" is the circuit's template parameters
module.exports = async function builder(options) { return new WitnessCalculator(options); };
module.exports.WitnessCalculator = WitnessCalculator;

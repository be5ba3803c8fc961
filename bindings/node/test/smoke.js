// node test/smoke.js <input.json> : RollupMain(3,16,2,2) witness of a generated input, hashGlobalInputs printed
// (input.json as tools/generate-input.js of the reference writes it, or `python -m circuits_b200.cli --dump-input`)
"use strict";
const fs = require("fs");
const builder = require("../witness_calculator.js");
(async () => {
    const input = JSON.parse(fs.readFileSync(process.argv[2], "utf8"));
    const wc = await builder({ circuit: "rollup-main", nTx: 3, nLevels: 16, maxL1Tx: 2, maxFeeTx: 2 });
    const w = await wc.calculateWitness(input, true);
    console.log("NVars", wc.NVars, "hashGlobalInputs", w[1].toString());
    await wc.writeWtns(w, "witness.wtns");
})().catch((e) => { console.error(e.message); process.exit(1); });

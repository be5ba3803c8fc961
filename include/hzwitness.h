/* hzwitness.h -- C ABI of the B200 witness generator for the Hermez circuits.
 *
 * This is the drop-in boundary: the entry points a Node N-API shim,
 * a Python ctypes binding or a C++ host would bind in place of the reference's witness
 * calculators.  Reference interfaces replaced, by entry point:
 *
 *   hzw_create / hzw_destroy      circom_runtime `builder(code, options)` that the reference's
 *                                 tests obtain through `tester(path, {reduceConstraints:false})`
 *                                 (/root/reference/test/rollup-main.test.js:52) and the compiled
 *                                 binary started at /root/reference/tools/helpers/actions.js:139;
 *                                 circuit parameters are the template parameters of
 *                                 /root/reference/src/rollup-main.circom:82 and withdraw.circom:21
 *   hzw_calculate                 `circuit.calculateWitness(input, opts)` /
 *                                 `calculateBinWitness`  (/root/reference/test/helpers/helpers.js:149,
 *                                 test/rollup-main.test.js:873) and `./circuit input.json witness.wtns`
 *                                 (/root/reference/tools/helpers/actions.js:139-143)
 *   hzw_input_index               the runtime's input-name lookup (`getSignalOffset32`), i.e. the keys
 *                                 of `bb.getInput()` (/root/reference/src/rollup-main.circom:105-196)
 *   hzw_get_signal / hzw_dump_sym circom tester `getSignal(w, name)` / `.sym` lookup behind
 *                                 `assertOut` (/root/reference/test/helpers/helpers.js:151-154,170)
 *   hzw_write_wtns                snarkjs `wtns calculate` output (`.wtns`, the witness file
 *                                 /root/reference/tools/helpers/actions.js:139 produces)
 *   hzw_calculate_batch           many independent witnesses of one circuit in a single call
 *                                 (withdraw, /root/reference/test/withdraw.test.js:105-160)
 *
 * All buffers are caller-owned.  Field elements cross the boundary as 32-byte little-endian
 * canonical residues of the BN254 scalar field.  A context is not re-entrant; distinct
 * contexts may be used from distinct threads.  Return value 0 = success, negative = error.
 */
#ifndef HZWITNESS_H
#define HZWITNESS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HZW_CIRCUIT_ROLLUP_MAIN 0
#define HZW_CIRCUIT_WITHDRAW 1
/* RollupTx(nLevels, maxFeeTx) as a circuit of its own (`component main = RollupTx(...)`, the way
 * /root/reference/test/rollup-tx.test.js:56-920 compiles it): inputs are the template's input signals in declaration
 * order (/root/reference/src/rollup-tx.circom:93-170); instances are independent (hzw_calculate / hzw_calculate_batch) */
#define HZW_CIRCUIT_ROLLUP_TX 2

#define HZW_OK 0
#define HZW_E_PARAM (-1)        /* bad parameters / null pointers */
#define HZW_E_IO (-2)           /* file could not be written */
#define HZW_E_NOSIGNAL (-3)     /* "Signal <name> is not an input of the circuit." / unknown signal */
#define HZW_E_CONSTRAINT (-4)   /* "Constraint doesn't match <lhs> != <rhs>" (sanity_check only) */
#define HZW_E_DEVICE (-5)       /* CUDA error; no CPU fallback exists */

typedef struct hzw_ctx hzw_ctx;

typedef struct {
    int32_t circuit_kind;       /* HZW_CIRCUIT_* */
    int32_t nTx, nLevels, maxL1Tx, maxFeeTx;   /* withdraw uses nLevels only */
    int32_t device;             /* CUDA device ordinal (the first one when n_gpus > 1) */
    int32_t n_gpus;             /* 0 or 1: one GPU.  N > 1: ONE context drives devices device .. device + N - 1 of this
                                 * box (they must reach each other peer to peer): the tx slots, fee txs and SHA blocks of
                                 * a batch are split over them, the exchange between the two phases is done by the pack
                                 * kernels storing into every peer over NVLink, and hzw_calculate lands ONE witness in
                                 * witness_out, each GPU copying the ranges it owns to their offsets over its own link */
    uint32_t flags;             /* reserved, 0 */
} hzw_params;

typedef struct {
    int32_t code;
    uint32_t line;              /* template source line of the failed `===` */
    uint64_t witness_index;     /* witness cursor where it failed: resolves to a component path */
    uint8_t lhs[32], rhs[32];
    char where[256];            /* component path when it could be resolved */
} hzw_error;

/* layout_map_path (nullable): a circom `.sym` of the same circuit (`circom -s`: lines `labelIdx,varIdx,componentIdx,name`).
 * With it every witness index that crosses this boundary is the COMPILER's: hzw_nvars() is the number of witness entries
 * of the compiler's layout (a reduced `-r` build keeps fewer than this library evaluates), hzw_calculate / hzw_write_wtns
 * deliver the entries in the compiler's order (a gather kernel on the device before the copy), hzw_signal_index /
 * hzw_get_signal / hzw_dump_sym answer in it.  Every name in the file must be a signal of this circuit and every witness
 * index must be named, else HZW_E_PARAM (the reason goes to stderr).  One GPU, HZW_CIRCUIT_ROLLUP_MAIN.  Replaces what
 * the reference gets from compiling with `sym: true` (/root/reference/test/rollup-main.test.js:52,
 * /root/reference/tools/helpers/actions.js:41). */
int hzw_create(const hzw_params* params, const char* layout_map_path /* nullable */, hzw_ctx** out);
void hzw_destroy(hzw_ctx* ctx);
uint64_t hzw_nvars(const hzw_ctx* ctx);
uint64_t hzw_ninputs(const hzw_ctx* ctx);
/* offset (in field elements, inside the packed input vector) and element count of a main input */
int hzw_input_index(const hzw_ctx* ctx, const char* name, uint64_t* offset, uint64_t* len);
/* inputs_le32: hzw_ninputs() x 32 B in declaration order; witness_out: hzw_nvars() x 32 B */
int hzw_calculate(hzw_ctx* ctx, const uint8_t* inputs_le32, int sanity_check, uint8_t* witness_out, hzw_error* err);
/* n independent instances: inputs n x ninputs x 32 B, witnesses n x nvars x 32 B (withdraw) */
int hzw_calculate_batch(hzw_ctx* ctx, uint64_t n, const uint8_t* inputs_le32, int sanity_check, uint8_t* witness_out,
                        hzw_error* err);
/* the batch with the witnesses left in device memory: public_out (nullable) receives witness entry 1 -- the public output
 * hashGlobalInputs -- of every instance, n x 32 B; kernel_ms = device time of the whole call (slowest GPU when n_gpus > 1:
 * the instances are split into contiguous ranges, one per GPU, no exchange).  Withdraw contexts only. */
int hzw_calculate_batch_resident(hzw_ctx* ctx, uint64_t n, const uint8_t* inputs_le32, int sanity_check, uint8_t* public_out,
                                 float* kernel_ms, hzw_error* err);
/* same computation, witness left in device memory (no device->host copy); kernel_ms = device time.
 * inputs_le32 == NULL re-uses the inputs already resident in device memory from the previous call. */
int hzw_calculate_resident(hzw_ctx* ctx, const uint8_t* inputs_le32, int sanity_check, float* kernel_ms, hzw_error* err);
/* ---- device-resident hand-off (the step after the path: /root/reference/tools/helpers/actions.js:148-185 feeds the
 * witness file to the prover) --------------------------------------------------------------------------------
 * After hzw_calculate_resident (or hzw_shard_phase2 with shard_out == NULL) the witness stays in HBM; a GPU prover
 * consumes it there instead of paying the PCIe egress:
 *   hzw_witness_device  base pointer of the witness in device memory (hzw_nvars() x 32 B, entry i at +32 i, little-
 *                       endian canonical residues, same layout as the host buffer of hzw_calculate), its CUDA device
 *                       ordinal and the stream (cudaStream_t) the producing kernels were ordered on; valid until the
 *                       next calculate on this context or hzw_destroy.  On a sharded context only the ranges of
 *                       hzw_shard_pieces hold this rank's results.
 *   hzw_signal_index    witness index of a named signal ("main.rollupTx[3].processor1.newRoot"): the layout query
 *   hzw_witness_read    device -> host copy of entries [first, first + count) of the resident witness */
int hzw_witness_device(const hzw_ctx* ctx, void** dev_ptr, int32_t* device, void** stream);
int hzw_signal_index(const hzw_ctx* ctx, const char* name, uint64_t* index);
int hzw_witness_read(hzw_ctx* ctx, uint64_t first, uint64_t count, uint8_t* out);
/* test introspection of the signature pipeline: `count` field elements from `first` of the work area of signature g
 * (which = 0: the hint scratch, 1: the points of its ladder walk, 2: the cached window table; see csrc/hints.cuh) */
int hzw_debug_read(hzw_ctx* ctx, int which, uint64_t g, uint64_t first, uint64_t count, uint8_t* out);
/* n_gpus > 1: after hzw_calculate_resident rank r's ranges (hzw_rank_pieces: (first index, count) pairs) are valid
 * in rank r's copy of the witness on its own device (hzw_rank_witness_device); hzw_witness_read gathers from them */
int hzw_rank_pieces(const hzw_ctx* ctx, int32_t rank, uint64_t* table /* 2 x cap */, int32_t cap);
int hzw_rank_witness_device(const hzw_ctx* ctx, int32_t rank, void** dev_ptr, int32_t* device, void** stream);
int hzw_write_wtns(const hzw_ctx* ctx, const uint8_t* witness, const char* path);
/* `./circuit input.json witness.wtns` in one call (/root/reference/tools/helpers/actions.js:139-143): the batch is
 * evaluated with the witness left in HBM and streamed into the file in chunks through two pinned buffers, each chunk
 * crossing PCIe while the previous one is written -- the 8 GB witness never sits in host memory.  With a layout map the
 * file holds the compiler's layout.  A violated constraint (sanity_check) writes no file.  HZW_CIRCUIT_ROLLUP_MAIN, one GPU. */
int hzw_calculate_to_wtns(hzw_ctx* ctx, const uint8_t* inputs_le32, int sanity_check, const char* path, float* kernel_ms /* nullable */,
                          hzw_error* err);
/* The constraints of an iden3 `.r1cs` (v1, BN254) evaluated over a witness: A.w * B.w = C.w, one GPU thread per constraint.
 * witness: host buffer of at least nWires entries in the numbering the file uses, or NULL = the witness resident on the
 * device after hzw_calculate_resident (gathered into the compiler's layout when the context has a layout map).
 * n_bad = violated constraints, first_bad = lowest index of one (~0 if none).  The independent check the reference's tests
 * get from `circuit.checkConstraints(witness)` / `loadConstraints()` (/root/reference/test/rollup-main.test.js:52-54). */
int hzw_r1cs_check(hzw_ctx* ctx, const char* r1cs_path, const uint8_t* witness /* nullable */, uint64_t* n_constraints,
                   uint64_t* n_bad, uint64_t* first_bad, float* kernel_ms /* nullable */);
int hzw_get_signal(const hzw_ctx* ctx, const uint8_t* witness, const char* name, uint8_t out[32]);
/* writes "index,name" lines (the `.sym` equivalent) for every signal whose name starts with prefix */
int hzw_dump_sym(const hzw_ctx* ctx, const char* prefix, const char* path);
/* sizes the layout walk derived, in witness entries: [0] nvars, [1] ninputs, [2] getAx, [3] sigVerifier,
 * [4] one sha256compression, [5] sha blocks, [6] feeTx[0], [7] decodeTx[1], [8] rollupTx[1] (0 if nTx < 3),
 * [9] processor1 + s3 + processor2 + s4 + s5 of a RollupTx, [10] offset of processor2 in it, [11] head of hasherInputs */
int hzw_plan_info(const hzw_ctx* ctx, uint64_t out[12]);
/* device time of each kernel of the last calculate, in launch order; names are static strings.
 * Returns the number of kernels (<= cap). */
int hzw_last_kernel_ms(const hzw_ctx* ctx, float* ms, const char** names, int cap);
/* One more batch on the inputs already resident, every launch serialised on one stream with an event in front of it:
 * device time of each launch ALONE, in launch order (a measurement aid for the roofline; the witness it leaves is the
 * same).  Returns the number of launches (<= cap). */
int hzw_profile_step(hzw_ctx* ctx, float* ms, const char** names, int cap);
/* out[0] = SMT processor levels this context evaluates per batch, out[1] = how many of them hashed (0, 0) on both paths
 * in the last batch (their entries are copies of one recorded image: k_smt_levels skips them, k_smt_zero copies them) */
int hzw_image_levels(hzw_ctx* ctx, uint64_t out[2]);
/* static description of what the last calculate launched: kernels and bytes, for the bench */
int hzw_last_stats(const hzw_ctx* ctx, uint64_t* gpu_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes);


/* ---- one batch on several GPUs -------------------------------------------------------------------
 * The tx slots of a batch are independent given the intermediary inputs (`im*`,
 * /root/reference/src/rollup-main.circom:93-99): the reference exploits this with one thread per
 * RollupTx/DecodeTx/FeeTx component (`circom -n`, /root/reference/tools/helpers/actions.js:41); here a
 * context per GPU evaluates a contiguous range of slots, fee txs and SHA blocks.  The only thing a
 * rank needs from the others is what hasherInputs hashes: their slots' pre-image bits and the three
 * chain ends (newLastIdx, newStateRoot, newExitRoot) -- one fixed-size record per rank, exchanged by
 * the caller with ONE all-gather (NCCL in the Python binding) between the two phases.
 *   hzw_set_shard      rank / world of this context (default 0 / 1: everything)
 *   hzw_exchange_bytes size of one rank's record
 *   hzw_shard_phase1   inputs -> own slots and fee txs; writes the own record to record_out.  With world > 1 it returns
 *                      as soon as the RECORD is complete (it is packed ahead of the merkle levels, which are still running:
 *                      the collective is meant to run beside them); a constraint failure of this rank's slots is then
 *                      reported by hzw_shard_phase2, which must follow
 *   hzw_shard_phase2   records = world x exchange_bytes, gathered; hasherInputs, own SHA blocks; copies the
 *                      witness pieces this rank owns to shard_out, back to back
 *   hzw_shard_pieces   the (witness index, count) pairs of those pieces; hzw_shard_nvars their total
 * record_out / records are DEVICE pointers for libhzwitness.so (the collective runs on device memory). */
int hzw_set_shard(hzw_ctx* ctx, int32_t rank, int32_t world);
uint64_t hzw_exchange_bytes(const hzw_ctx* ctx);
int hzw_shard_pieces(const hzw_ctx* ctx, uint64_t* table /* 2 x cap */, int32_t cap);
uint64_t hzw_shard_nvars(const hzw_ctx* ctx);
int hzw_shard_phase1(hzw_ctx* ctx, const uint8_t* inputs_le32, int sanity_check, void* record_out, hzw_error* err);
int hzw_shard_phase2(hzw_ctx* ctx, const void* records, int sanity_check, uint8_t* shard_out, float* kernel_ms, hzw_error* err);

/* Poseidon alone, the second metric of the benchmark: `npairs` lane groups, each running two
 * Poseidon(2) permutations (the old-path / new-path hash pair of an SMT level,
 * circomlib smtprocessorlevel.circom via /root/reference/src/rollup-tx.circom:537) on the given
 * inputs and storing all signals of both components, |reps| times; device ms per repetition.  reps > 0 runs the
 * transposed kernel (thread = one state element of one permutation, 32 permutations per 3-warp block), reps < 0 the
 * lane-group form used inside the SMT kernel; both store the same signals at the same indexes. */
int hzw_poseidon_bench(int device, uint64_t npairs, int reps, const uint8_t* inputs_le32 /* 4 x npairs x 32 B */,
                       uint8_t* digests_le32 /* 2 x npairs x 32 B, nullable */, float* ms_per_rep);
/* the same for the other widths of the path (t = 4 leaf hashes, 5 state hashes, 6 the verifier's hash, 7 the message
 * hash): nperms Poseidon(t - 1) hashes, transposed kernel, every signal of the component stored; digests nullable */
int hzw_poseidon_bench_t(int device, int t, uint64_t nperms, int reps, const uint8_t* inputs_le32 /* (t - 1) x nperms x 32 B */,
                         uint8_t* digests_le32 /* nperms x 32 B */, float* ms_per_rep, uint64_t* signal_bytes);
/* digests only: n independent Poseidon(t - 1) hashes (t = 3..7) on the device, nothing stored but the result.  What the
 * batch builder's trees hash with, one call per wave (circuits_b200/batchtree.py; the reference builds its batches with
 * circomlib's smt.js, /root/reference/tools/generate-input.js:61-109). */
int hzw_poseidon_hash(int device, int t, uint64_t n, const uint8_t* inputs_le32 /* (t - 1) x n x 32 B */, uint8_t* digests_le32 /* n x 32 B */);

/* Measured ceiling of the integer-multiply pipe: 32 x 32 -> 64-bit multiply-adds per second over all
 * lanes of the device (independent chains, full occupancy); the denominator of bench.py's integer roofline. */
int hzw_int_peak(int device, double* lane_products_per_s);

#ifdef __cplusplus
}
#endif
#endif

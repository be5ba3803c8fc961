// Interface between the host half of the library (hzw_host.cpp: C ABI, layout walk) and the
// CUDA half (hzw_kernels.cu: device memory, kernels, streams).
#pragma once
#include "rollup_main.cuh"
#include "withdraw.cuh"
#include "rollup_tx.cuh"

struct HzCuda;   // device state of one context
// returns null and sets *status (HZW_E_DEVICE) when no usable CUDA device / allocation failure
HzCuda* hz_cuda_create(int device, const HzPlan& plan_host, const uint32_t* off_decode, const uint32_t* off_rollup,
                       const uint32_t* off_fee, const uint32_t* off_eddsa, const uint32_t* off_jk,
                       const uint32_t* off_hs /* 4 nTx + 2 maxFeeTx */, const uint32_t* off_hsig /* nTx */, int* status);
void hz_cuda_destroy(HzCuda*);
// inputs: n_inputs x 32 B host; witness_out: nvars x 32 B host or null (resident)
int hz_cuda_calculate(HzCuda*, const uint8_t* inputs, int sanity, uint8_t* witness_out, float* kernel_ms,
                      HzErrBuf* errs, uint64_t* launches, uint64_t* h2d, uint64_t* d2h);
// device-resident hand-off: base of the witness in HBM, its device and the stream its kernels were ordered on
int hz_cuda_witness_device(HzCuda*, void** ptr, int32_t* device, void** stream);
int hz_cuda_witness_read(HzCuda*, uint64_t first, uint64_t count, uint8_t* out);
int hz_cuda_debug_read(HzCuda*, int which, uint64_t g, uint64_t first, uint64_t count, uint8_t* out);
// per-kernel device times of the last calculate
int hz_cuda_kernel_ms(HzCuda*, float* ms, const char** names, int cap);
// one more batch with every launch serialised on one stream: device time of each launch alone, in launch order
int hz_cuda_profile_step(HzCuda*, float* ms, const char** names, int cap);
int hz_cuda_image_levels(HzCuda*, uint64_t out[2]);

// ---- sharded evaluation (one context per GPU; include/hzwitness.h, hzw_shard_*) -------------------
int hz_cuda_set_shard(HzCuda*, int rank, int world);
uint64_t hz_cuda_exchange_bytes(const HzCuda*);
uint64_t hz_cuda_launches(const HzCuda*);     // kernels launched since the start of the current calculate / shard phase 1
int hz_cuda_shard_phase1(HzCuda*, const uint8_t* inputs, int sanity, void* record_out_dev, HzErrBuf* errs);
// in_place: every piece goes to its own place in a whole-witness buffer instead of back to back into a shard buffer
int hz_cuda_shard_phase2(HzCuda*, const void* records_dev, int sanity, uint8_t* shard_out, const uint64_t* pieces, int npieces,
                         float* kernel_ms, HzErrBuf* errs, uint64_t* d2h, bool in_place = false);
// one process driving several GPUs: peer access, phase 1 whose pack kernel stores the record into every peer, the wait
int hz_cuda_enable_peer(HzCuda*, HzCuda* peer);
uint8_t* hz_cuda_recs(HzCuda*);
int hz_cuda_multi_phase1(HzCuda*, const uint8_t* inputs, int sanity, uint8_t* const* peer_recs, int world);
int hz_cuda_multi_wait(HzCuda*, HzCuda* peer);

// ---- formats around the path (k_formats.cu): layout map of a circom .sym, .r1cs constraint check ----------------
struct HzR1cs;
int hz_cuda_set_map(HzCuda*, const uint32_t* src, uint64_t n_out);
int hz_cuda_remap(HzCuda*);                  // resident witness -> the compiler's layout, on the main stream
int hz_cuda_read_mapped(HzCuda*, uint64_t first, uint64_t count, uint8_t* out);
// the resident witness (gathered one under a layout map) written behind `header` in chunks, copies and writes overlapped
int hz_cuda_stream_wtns(HzCuda*, const char* path, const uint8_t* header, size_t header_len, uint64_t* d2h);
int hz_cuda_r1cs_check(HzCuda*, const HzR1cs& R, const uint8_t* witness, uint64_t* n_bad, uint64_t* first_bad, float* ms);

// ---- Withdraw(nLevels) batches ---------------------------------------------------------------------
int hz_cuda_probe(int device);       // 0 when the device exists, else HZW_E_DEVICE / HZW_E_PARAM
struct HzWdCuda;                     // device state of a Withdraw context (two witness buffers, two streams)
// maxFeeTx == 0: Withdraw(nLevels); > 0: the standalone RollupTx(nLevels, maxFeeTx) -- both are batches of independent instances
HzWdCuda* hz_cuda_withdraw_create(int device, int nLevels, int maxFeeTx, uint32_t n_inputs, uint32_t nvars, int* status);
void hz_cuda_withdraw_destroy(HzWdCuda*);
// n instances: inputs n x ninputs x 32 B; out (nullable) n x nvars x 32 B; pub_out (nullable) n x 32 B = witness entry 1
// of every instance; kernel_ms = device time of the whole call (chunks of the batch pipelined on two streams)
int hz_cuda_withdraw_batch(HzWdCuda*, uint64_t n, const uint8_t* inputs, int sanity, uint8_t* out, uint8_t* pub_out,
                           float* kernel_ms, HzErrBuf* errs, uint64_t* launches);

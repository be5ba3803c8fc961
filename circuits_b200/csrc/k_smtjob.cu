// The SMT job records of a slot ahead of k_front (see hzw_dev.cuh for the map of the kernels).
#include "hzw_dev.cuh"

// ---- the SMT jobs of a slot ahead of k_front --------------------------------------------------------------
// The merkle kernels are the other long chain of a slot, and what they need from rollupTx[i] are the two job
// records (keys, old / new state hashes, function selectors).  This kernel runs the SAME templates as k_front on
// a dry context -- values only: no entry stored, no inverse queued, the signature and the fee accumulator
// skipped, decodeTx reduced to the fields rollupTx reads -- so the merkle kernels start after a fraction of
// k_front's time, and k_front itself runs beside them.  k_front still records the jobs it arrives at; k_tail
// compares the two (always: a difference is an internal error, HzErrBuf.pad).
__global__ void __launch_bounds__(32) k_smtjob(HzPlan pl, Fr* w, HzSmtJob* smt, Fr* isExit, const uint32_t* off_jk, Fr* hs_in, int s0, int ns) {
    int g = (int)(blockIdx.x * blockDim.x + threadIdx.x);           // one thread per slot
    if (g >= ns) g = ns - 1;
    const int i = s0 + g;
    Comp<MCompute> c(w, pl.off_rollup[i], nullptr);
    c.dry = true; c.solo = true;
    int hs_n = 0;
    c.skip_hash = true; c.hs_rec = hs_in + 16 * (size_t)i; c.hs_n = &hs_n;       // the four state hashes: inputs recorded for k_hstate
    DecodeOut<MCompute> d;
    hz_decode_lite(c, pl, i, d);
    c.cur = pl.off_rollup[i];                 // the decode fields were evaluated on this context: back to where rollupTx[i] starts
    PreBits<MCompute> pre; pre.b = nullptr;
    SlotHook<MCompute> hook; hook.eddsa = nullptr; hook.smt = &smt[2 * i]; hook.isExit = &isExit[i]; hook.sz = &pl.sz; hook.jobs_only = true;
    RollupTxOut<MCompute> o;
    sec_rollup(c, pl, i, d, o, pre, hook);
    // the dry walk skips the fee accumulator: the processors start where the layout walk put them
    smt[2 * i].cur = off_jk[i]; smt[2 * i + 1].cur = off_jk[i] + pl.sz.p2_off;
}

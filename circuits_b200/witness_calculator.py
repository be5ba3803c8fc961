"""Host-side mirror of the witness-calculator surface the reference drives its circuits with.

`WitnessCalculator` reproduces the interface of circom_runtime 0.1.12 `WitnessCalculator`
(the object `tester(path)` / `builder(wasm)` hand to the reference's tests,
/root/reference/test/rollup-main.test.js:52, test/helpers/helpers.js:142-155):
`calculateWitness(input, sanityCheck) -> list[int]`, `calculateBinWitness(input, sanityCheck)
-> bytes` (NVars x 32 B little-endian), plus the tester helpers `assertOut`, `getSignal` and a
`.wtns` writer (snarkjs `wtns calculate`).  Errors follow the runtime's wording: a violated
`===` raises `WitnessError("Constraint doesn't match <lhs> != <rhs> ...")`, an unknown key
`"Signal <k> is not an input of the circuit."`.

All arithmetic happens in `libhzwitness.so` (hand-written CUDA, sm_100a) through the C ABI of
`include/hzwitness.h`; there is no CPU path -- construction fails when the library or a CUDA
device is missing.
"""
import ctypes
import os

from .field import P

HZW_CIRCUIT_ROLLUP_MAIN = 0
HZW_CIRCUIT_WITHDRAW = 1
HZW_CIRCUIT_ROLLUP_TX = 2
_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libhzwitness.so")


class WitnessError(Exception):
    """Mirrors the `Error` circom_runtime throws."""


class _Params(ctypes.Structure):
    _fields_ = [("circuit_kind", ctypes.c_int32), ("nTx", ctypes.c_int32), ("nLevels", ctypes.c_int32),
                ("maxL1Tx", ctypes.c_int32), ("maxFeeTx", ctypes.c_int32), ("device", ctypes.c_int32),
                ("n_gpus", ctypes.c_int32), ("flags", ctypes.c_uint32)]


class _Error(ctypes.Structure):
    _fields_ = [("code", ctypes.c_int32), ("line", ctypes.c_uint32), ("witness_index", ctypes.c_uint64),
                ("lhs", ctypes.c_uint8 * 32), ("rhs", ctypes.c_uint8 * 32), ("where", ctypes.c_char * 256)]


EXPORTS = ["hzw_create", "hzw_destroy", "hzw_nvars", "hzw_ninputs", "hzw_input_index", "hzw_calculate",
           "hzw_calculate_batch", "hzw_calculate_resident", "hzw_write_wtns", "hzw_get_signal", "hzw_dump_sym",
           "hzw_last_stats", "hzw_last_kernel_ms", "hzw_plan_info", "hzw_set_shard", "hzw_exchange_bytes",
           "hzw_shard_pieces", "hzw_shard_nvars", "hzw_shard_phase1", "hzw_shard_phase2", "hzw_poseidon_bench", "hzw_int_peak",
           "hzw_witness_device", "hzw_signal_index", "hzw_witness_read", "hzw_profile_step", "hzw_image_levels", "hzw_rank_pieces", "hzw_rank_witness_device", "hzw_calculate_batch_resident", "hzw_r1cs_check", "hzw_calculate_to_wtns", "hzw_debug_read", "hzw_poseidon_bench_t", "hzw_poseidon_hash"]


def load_library(path=None):
    path = path or DEFAULT_LIB
    if not os.path.exists(path):
        raise WitnessError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    lib.hzw_nvars.restype = ctypes.c_uint64
    lib.hzw_ninputs.restype = ctypes.c_uint64
    lib.hzw_create.argtypes = [ctypes.POINTER(_Params), ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
    lib.hzw_destroy.argtypes = [ctypes.c_void_p]
    lib.hzw_nvars.argtypes = [ctypes.c_void_p]
    lib.hzw_ninputs.argtypes = [ctypes.c_void_p]
    lib.hzw_input_index.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64),
                                    ctypes.POINTER(ctypes.c_uint64)]
    lib.hzw_calculate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.POINTER(_Error)]
    lib.hzw_calculate_batch.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int,
                                        ctypes.c_void_p, ctypes.POINTER(_Error)]
    lib.hzw_calculate_resident.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                           ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_Error)]
    lib.hzw_write_wtns.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p]
    lib.hzw_get_signal.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p]
    lib.hzw_dump_sym.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p]
    lib.hzw_last_kernel_ms.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float),
                                       ctypes.POINTER(ctypes.c_char_p), ctypes.c_int]
    lib.hzw_profile_step.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float),
                                     ctypes.POINTER(ctypes.c_char_p), ctypes.c_int]
    lib.hzw_plan_info.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
    lib.hzw_last_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64),
                                   ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64)]
    lib.hzw_set_shard.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32]
    lib.hzw_exchange_bytes.argtypes = [ctypes.c_void_p]
    lib.hzw_exchange_bytes.restype = ctypes.c_uint64
    lib.hzw_shard_pieces.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64), ctypes.c_int32]
    lib.hzw_shard_nvars.argtypes = [ctypes.c_void_p]
    lib.hzw_shard_nvars.restype = ctypes.c_uint64
    lib.hzw_shard_phase1.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.POINTER(_Error)]
    lib.hzw_shard_phase2.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                     ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_Error)]
    lib.hzw_witness_device.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int32),
                                       ctypes.POINTER(ctypes.c_void_p)]
    lib.hzw_signal_index.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint64)]
    lib.hzw_witness_read.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
    return lib


def _flatten(v, out):
    if isinstance(v, (list, tuple)):
        for x in v:
            _flatten(x, out)
    else:
        out.append(int(v) % P)


def _addr(buf):
    """Address of a writable/readable buffer (bytearray, ctypes array, numpy / torch storage)."""
    if isinstance(buf, int):
        return buf
    if hasattr(buf, "data_ptr"):
        return buf.data_ptr()
    if hasattr(buf, "ctypes"):
        return buf.ctypes.data
    return ctypes.addressof((ctypes.c_uint8 * len(buf)).from_buffer(buf))


class WitnessCalculator:
    prime = P
    n64 = 4

    def __init__(self, nTx, nLevels, maxL1Tx, maxFeeTx, circuit="rollup-main", device=0, lib_path=None, n_gpus=1, sym=None):
        self._lib = load_library(lib_path)
        kind = {"rollup-main": HZW_CIRCUIT_ROLLUP_MAIN, "withdraw": HZW_CIRCUIT_WITHDRAW, "rollup-tx": HZW_CIRCUIT_ROLLUP_TX}[circuit]
        self.params = (nTx, nLevels, maxL1Tx, maxFeeTx)
        prm = _Params(kind, nTx, nLevels, maxL1Tx, maxFeeTx, device, n_gpus, 0)
        self._ctx = ctypes.c_void_p()
        # sym: path of a circom `.sym` of this circuit -- witness indexes then follow the compiler's layout
        st = self._lib.hzw_create(ctypes.byref(prm), None if sym is None else str(sym).encode(), ctypes.byref(self._ctx))
        if st != 0:
            self._ctx = None
            raise WitnessError("hzw_create failed with code %d (no CUDA device? bad parameters?)" % st)
        self.NVars = int(self._lib.hzw_nvars(self._ctx))
        self.NInputs = int(self._lib.hzw_ninputs(self._ctx))

    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.hzw_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- input packing: what circom_runtime's `setSignal` loop does ------------------------
    def inputIndex(self, name):
        off, ln = ctypes.c_uint64(), ctypes.c_uint64()
        if self._lib.hzw_input_index(self._ctx, name.encode(), ctypes.byref(off), ctypes.byref(ln)) != 0:
            raise WitnessError("Signal %s is not an input of the circuit." % name)
        return off.value, ln.value

    def packInput(self, inp, out=None):
        """Flatten an input object (keys = main input names) into NInputs x 32 B."""
        buf = out if out is not None else bytearray(32 * self.NInputs)
        mv = memoryview(buf)
        seen = 0
        for name, v in inp.items():
            off, ln = self.inputIndex(name)
            vals = []
            _flatten(v, vals)
            if len(vals) != ln:
                raise WitnessError("Invalid number of values for input signal %s: %d != %d" % (name, len(vals), ln))
            mv[32 * off:32 * (off + ln)] = b"".join(x.to_bytes(32, "little") for x in vals)
            seen += ln
        if seen != self.NInputs:
            raise WitnessError("Not all inputs have been set. Only %d out of %d" % (seen, self.NInputs))
        return buf

    def _raise(self, st, err):
        if st == -4:
            lhs = int.from_bytes(bytes(err.lhs), "little")
            rhs = int.from_bytes(bytes(err.rhs), "little")
            raise WitnessError("Constraint doesn't match %d != %d %s (template line %d)"
                               % (lhs, rhs, err.where.decode(), err.line))
        raise WitnessError("hzw_calculate failed with code %d" % st)

    # -- circom_runtime surface ----------------------------------------------------------------
    def calculateBinWitness(self, inp, sanityCheck=False, out=None):
        packed = inp if isinstance(inp, (bytes, bytearray)) or hasattr(inp, "data_ptr") or hasattr(inp, "ctypes") else self.packInput(inp)
        if isinstance(packed, bytes):
            packed = bytearray(packed)
        wit = out if out is not None else bytearray(32 * self.NVars)
        err = _Error()
        st = self._lib.hzw_calculate(self._ctx, _addr(packed), 1 if sanityCheck else 0, _addr(wit), ctypes.byref(err))
        if st != 0:
            self._raise(st, err)
        return wit

    def calculateToWtns(self, inp, path, sanityCheck=False):
        """`./circuit input.json witness.wtns` in one call (`hzw_calculate_to_wtns`): the witness is streamed from
        device memory into the file; returns the device ms of the kernels."""
        packed = inp if isinstance(inp, (bytes, bytearray)) or hasattr(inp, "data_ptr") or hasattr(inp, "ctypes") else self.packInput(inp)
        if isinstance(packed, bytes):
            packed = bytearray(packed)
        err, ms = _Error(), ctypes.c_float()
        self._lib.hzw_calculate_to_wtns.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p,
                                                    ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_Error)]
        st = self._lib.hzw_calculate_to_wtns(self._ctx, _addr(packed), 1 if sanityCheck else 0, str(path).encode(),
                                             ctypes.byref(ms), ctypes.byref(err))
        if st != 0:
            self._raise(st, err)
        return ms.value

    def calculateBatch(self, n, packed, sanityCheck=False, out=None):
        """`n` independent witnesses of one circuit in a single call (`hzw_calculate_batch`; Withdraw)."""
        if isinstance(packed, bytes):
            packed = bytearray(packed)
        wit = out if out is not None else bytearray(32 * self.NVars * n)
        err = _Error()
        st = self._lib.hzw_calculate_batch(self._ctx, n, _addr(packed), 1 if sanityCheck else 0, _addr(wit), ctypes.byref(err))
        if st != 0:
            self._raise(st, err)
        return wit

    def calculateBatchResident(self, n, packed, sanityCheck=False, public_out=None):
        """`n` independent witnesses left in device memory (`hzw_calculate_batch_resident`; Withdraw): returns
        (the public output of every instance, n x 32 B; device ms of the call)."""
        if isinstance(packed, bytes):
            packed = bytearray(packed)
        pub = public_out if public_out is not None else bytearray(32 * n)
        err = _Error()
        ms = ctypes.c_float()
        self._lib.hzw_calculate_batch_resident.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_int,
                                                           ctypes.c_void_p, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_Error)]
        st = self._lib.hzw_calculate_batch_resident(self._ctx, n, _addr(packed), 1 if sanityCheck else 0, _addr(pub),
                                                    ctypes.byref(ms), ctypes.byref(err))
        if st != 0:
            self._raise(st, err)
        return pub, ms.value

    def calculateWitness(self, inp, sanityCheck=False):
        b = bytes(self.calculateBinWitness(inp, sanityCheck))
        return [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(self.NVars)]

    def calculateResident(self, packed, sanityCheck=False):
        """Witness left in device memory; returns the device time of the kernels in ms."""
        ms = ctypes.c_float()
        err = _Error()
        st = self._lib.hzw_calculate_resident(self._ctx, None if packed is None else _addr(packed),
                                              1 if sanityCheck else 0, ctypes.byref(ms),
                                              ctypes.byref(err))
        if st != 0:
            self._raise(st, err)
        return ms.value

    # -- device-resident hand-off ------------------------------------------------------------------
    def witnessDevice(self):
        """(device pointer, CUDA device ordinal, cudaStream_t) of the witness a resident calculate left in HBM."""
        ptr, dev, st = ctypes.c_void_p(), ctypes.c_int32(), ctypes.c_void_p()
        if self._lib.hzw_witness_device(self._ctx, ctypes.byref(ptr), ctypes.byref(dev), ctypes.byref(st)) != 0:
            raise WitnessError("this backend keeps no device-resident witness")
        return ptr.value, dev.value, st.value

    def signalIndex(self, name):
        idx = ctypes.c_uint64()
        if self._lib.hzw_signal_index(self._ctx, name.encode(), ctypes.byref(idx)) != 0:
            raise WitnessError("Signal %s not found" % name)
        return idx.value

    def readWitness(self, first=0, count=None, out=None):
        """Entries [first, first + count) of the resident witness, copied to the host."""
        count = self.NVars - first if count is None else count
        buf = out if out is not None else bytearray(32 * count)
        if self._lib.hzw_witness_read(self._ctx, first, count, _addr(buf)) != 0:
            raise WitnessError("hzw_witness_read(%d, %d) failed" % (first, count))
        return buf

    def checkConstraints(self, r1cs_path, witness=None):
        """`circuit.checkConstraints(witness)` of the circom tester on the GPU: every constraint of an iden3 .r1cs
        over `witness` (bytes in the file's wire numbering) or, with None, over the witness resident on the device.
        Returns (constraints, violated, index of the first violated one or None, device ms)."""
        n, bad, first, ms = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_float()
        self._lib.hzw_r1cs_check.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64),
                                             ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_float)]
        wb = None if witness is None else (witness if hasattr(witness, "ctypes") or hasattr(witness, "data_ptr") else bytearray(witness))
        st = self._lib.hzw_r1cs_check(self._ctx, str(r1cs_path).encode(), None if wb is None else _addr(wb), ctypes.byref(n),
                                      ctypes.byref(bad), ctypes.byref(first), ctypes.byref(ms))
        if st != 0:
            raise WitnessError("hzw_r1cs_check(%s) failed with code %d" % (r1cs_path, st))
        return n.value, bad.value, (None if bad.value == 0 else first.value), ms.value

    def planInfo(self):
        out = (ctypes.c_uint64 * 12)()
        self._lib.hzw_plan_info(self._ctx, out)
        keys = ["nvars", "ninputs", "getAx", "sigVerifier", "sha_block", "sha_blocks", "feeTx", "decodeTx", "rollupTx",
                "processors", "processor2_off", "hasher_head"]
        return dict(zip(keys, [int(x) for x in out]))

    def lastKernelMs(self):
        """[(kernel name, device ms)] of the last calculate, in launch order."""
        ms = (ctypes.c_float * 16)()
        names = (ctypes.c_char_p * 16)()
        n = self._lib.hzw_last_kernel_ms(self._ctx, ms, names, 16)
        return [(names[i].decode(), ms[i]) for i in range(max(n, 0))]

    def profileStep(self):
        """[(kernel name, device ms)] of one more batch on the resident inputs with every launch serialised on one
        stream (`hzw_profile_step`): the time each launch takes ALONE."""
        ms = (ctypes.c_float * 48)()
        names = (ctypes.c_char_p * 48)()
        n = self._lib.hzw_profile_step(self._ctx, ms, names, 48)
        if n < 0:
            raise WitnessError("hzw_profile_step failed with code %d" % n)
        return [(names[i].decode(), ms[i]) for i in range(n)]

    def imageLevels(self):
        """(SMT processor levels per batch, how many of them were (0, 0) images in the last batch)."""
        out = (ctypes.c_uint64 * 2)()
        self._lib.hzw_image_levels.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        if self._lib.hzw_image_levels(self._ctx, out) != 0:
            raise WitnessError("hzw_image_levels failed")
        return int(out[0]), int(out[1])

    def lastStats(self):
        a, b, c = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
        self._lib.hzw_last_stats(self._ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return dict(gpu_launches=a.value, h2d_bytes=b.value, d2h_bytes=c.value)

    # -- circom tester helpers (test/helpers/helpers.js:151-154,170) ------------------------------
    def getSignal(self, witness, name):
        wb = witness if isinstance(witness, (bytes, bytearray)) else b"".join(int(x).to_bytes(32, "little") for x in witness)
        wb = bytearray(wb)
        out = (ctypes.c_uint8 * 32)()
        if self._lib.hzw_get_signal(self._ctx, _addr(wb), name.encode(), out) != 0:
            raise WitnessError("Signal %s not found" % name)
        return int.from_bytes(bytes(out), "little")

    def assertOut(self, witness, expected):
        for k, v in expected.items():
            vals = []
            _flatten(v, vals)
            if isinstance(v, (list, tuple)):
                for i, e in enumerate(vals):
                    got = self.getSignal(witness, "main.%s[%d]" % (k, i))
                    if got != e:
                        raise AssertionError("main.%s[%d]: %d != %d" % (k, i, got, e))
            else:
                got = self.getSignal(witness, "main." + k)
                if got != vals[0]:
                    raise AssertionError("main.%s: %d != %d" % (k, got, vals[0]))

    def symbols(self, prefix="main"):
        """name -> witness index for every signal under `prefix` (the `.sym` equivalent)."""
        import tempfile
        with tempfile.NamedTemporaryFile(suffix=".sym", delete=False) as f:
            path = f.name
        try:
            if self._lib.hzw_dump_sym(self._ctx, prefix.encode(), path.encode()) != 0:
                raise WitnessError("hzw_dump_sym failed")
            table = {}
            with open(path) as fh:
                for line in fh:
                    i, n = line.rstrip("\n").split(",", 1)
                    table[n] = int(i)
            return table
        finally:
            os.unlink(path)

    def writeWtns(self, witness, path):
        wb = bytearray(witness)
        if self._lib.hzw_write_wtns(self._ctx, _addr(wb), path.encode()) != 0:
            raise WitnessError("cannot write %s" % path)


class ShardedWitnessCalculator(WitnessCalculator):
    """One batch on `world` GPUs, one process per GPU (`torch.distributed`).

    The circuit cuts every cross-slot dependency with its `im*` inputs
    (/root/reference/src/rollup-main.circom:93-99; the reference runs a thread per RollupTx /
    DecodeTx / FeeTx component on that basis, tools/helpers/actions.js:41).  Rank r evaluates a
    contiguous range of tx slots, fee txs and SHA blocks; what hasherInputs needs from the other
    ranks (their slots' pre-image bits and the three chain ends) travels in ONE all-gather between
    the two phases.  Each rank ends up with the witness pieces it owns (`pieces()`), copied back to
    back into its own host buffer -- the 8 GB witness never crosses NVLink.

    `record_device`: "cuda" for libhzwitness.so (records are device memory, collective = NCCL),
    "cpu" for the CPU port used by the gloo tests.
    """

    def __init__(self, nTx, nLevels, maxL1Tx, maxFeeTx, rank, world, device=0, lib_path=None, group=None,
                 record_device="cuda"):
        super().__init__(nTx, nLevels, maxL1Tx, maxFeeTx, device=device, lib_path=lib_path)
        import torch
        self.rank, self.world, self.group = rank, world, group
        if self._lib.hzw_set_shard(self._ctx, rank, world) != 0:
            raise WitnessError("hzw_set_shard(%d, %d) failed" % (rank, world))
        self.recBytes = int(self._lib.hzw_exchange_bytes(self._ctx))
        dev = torch.device("cuda", device) if record_device == "cuda" else torch.device("cpu")
        self._rec = torch.zeros(self.recBytes, dtype=torch.uint8, device=dev)
        self._recs = torch.zeros(self.recBytes * world, dtype=torch.uint8, device=dev)
        self.shardNVars = int(self._lib.hzw_shard_nvars(self._ctx))

    def pieces(self):
        """[(witness index, count)] of the pieces this rank owns, in the order they are copied out."""
        t = (ctypes.c_uint64 * 32)()
        n = self._lib.hzw_shard_pieces(self._ctx, t, 16)
        return [(int(t[2 * k]), int(t[2 * k + 1])) for k in range(n)]

    def phase1(self, packed, sanityCheck=False):
        """Own slots and fee txs; leaves this rank's exchange record in `self._rec`."""
        err = _Error()
        st = self._lib.hzw_shard_phase1(self._ctx, None if packed is None else _addr(packed), 1 if sanityCheck else 0,
                                        self._rec.data_ptr(), ctypes.byref(err))
        if st != 0:
            self._raise(st, err)

    def phase2(self, sanityCheck=False, out=None, resident=False):
        """hasherInputs from the gathered records in `self._recs`; copies the own pieces out."""
        err = _Error()
        ms = ctypes.c_float()
        buf = None if resident else (out if out is not None else bytearray(32 * self.shardNVars))
        st = self._lib.hzw_shard_phase2(self._ctx, self._recs.data_ptr(), 1 if sanityCheck else 0,
                                        None if buf is None else _addr(buf), ctypes.byref(ms), ctypes.byref(err))
        if st != 0:
            self._raise(st, err)
        return buf, ms.value

    def exchange(self):
        """The one collective of a batch: all-gather of the per-rank records."""
        import torch
        import torch.distributed as dist
        if self.world > 1:
            if self._rec.is_cuda:
                # on a stream of its own: torch's default stream is the legacy NULL stream, and a collective queued behind
                # it would wait for every kernel of the library's (blocking) streams -- phase 1 returns while the merkle
                # levels are still running (early exchange), and the all-gather is to run beside them
                if getattr(self, "_xs", None) is None:
                    self._xs = torch.cuda.Stream(device=self._rec.device)
                with torch.cuda.stream(self._xs):
                    dist.all_gather_into_tensor(self._recs, self._rec, group=self.group)
                self._xs.synchronize()
            else:
                parts = list(self._recs.split(self.recBytes))
                dist.all_gather(parts, self._rec, group=self.group)
        else:
            self._recs.copy_(self._rec)
            if self._rec.is_cuda:
                torch.cuda.current_stream().synchronize()

    def calculateShard(self, packed, sanityCheck=False, out=None, resident=False):
        """Both phases and the all-gather; returns (shard buffer, device ms of this rank's kernels).
        A rank whose phase 1 fails (a violated constraint under sanityCheck is a per-rank event: only the rank
        that owns the slot sees it) still takes part in the collective, so the other ranks are not left
        waiting in the all-gather; it raises afterwards."""
        failure = None
        try:
            self.phase1(packed, sanityCheck)
        except WitnessError as e:
            failure = e
        self.exchange()
        if failure is not None:
            raise failure
        return self.phase2(sanityCheck, out, resident)

    @staticmethod
    def assemble(nvars, shards):
        """Full witness from [(pieces, shard bytes)] of all ranks (tests / single-host consumers)."""
        w = bytearray(32 * nvars)
        for pcs, data in shards:
            off = 0
            mv = memoryview(data)
            for idx, cnt in pcs:
                w[32 * idx:32 * (idx + cnt)] = mv[off:off + 32 * cnt]
                off += 32 * cnt
        return w


def poseidon_pairs(inputs, reps=1, device=0, lib_path=None, lanes=False):
    """Run Poseidon(2) on pairs of pairs of field elements on the GPU (`hzw_poseidon_bench`):
    inputs = [(a, b, c, d), ...] -> ([(poseidon([a, b]), poseidon([c, d])), ...], device ms per repetition).
    lanes=True uses the lane-group kernel (the form inside k_smt), else the transposed one."""
    lib = load_library(lib_path)
    lib.hzw_poseidon_bench.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.POINTER(ctypes.c_float)]
    n = len(inputs)
    buf = bytearray(b"".join(int(x % P).to_bytes(32, "little") for quad in inputs for x in quad))
    out = bytearray(64 * n)
    ms = ctypes.c_float()
    st = lib.hzw_poseidon_bench(device, n, -reps if lanes else reps, _addr(buf), _addr(out), ctypes.byref(ms))
    if st != 0:
        raise WitnessError("hzw_poseidon_bench failed with code %d" % st)
    dig = [(int.from_bytes(out[64 * i:64 * i + 32], "little"), int.from_bytes(out[64 * i + 32:64 * i + 64], "little"))
           for i in range(n)]
    return dig, ms.value


def int_peak(device=0, lib_path=None):
    """Measured 32x32->64 multiply-adds per second of the device (all lanes), `hzw_int_peak`."""
    lib = load_library(lib_path)
    lib.hzw_int_peak.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    v = ctypes.c_double()
    if lib.hzw_int_peak(device, ctypes.byref(v)) != 0:
        raise WitnessError("hzw_int_peak failed")
    return v.value


def poseidon_width(t, inputs, reps=1, device=0):
    """Poseidon(t - 1) of every tuple in `inputs` on the GPU with the transposed kernel of width t = 4..7
    (`hzw_poseidon_bench_t`, every signal of the component stored): (digests, ms per launch, signal bytes per launch)."""
    lib = load_library()
    lib.hzw_poseidon_bench_t.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                         ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_uint64)]
    n = len(inputs)
    buf = bytearray(b"".join(int(x).to_bytes(32, "little") for tup in inputs for x in tup))
    assert len(buf) == 32 * (t - 1) * n
    out = bytearray(32 * n)
    ms, sb = ctypes.c_float(), ctypes.c_uint64()
    st = lib.hzw_poseidon_bench_t(device, t, n, reps, _addr(buf), _addr(out), ctypes.byref(ms), ctypes.byref(sb))
    if st != 0:
        raise WitnessError("hzw_poseidon_bench_t failed with code %d" % st)
    return [int.from_bytes(out[32 * i:32 * i + 32], "little") for i in range(n)], ms.value, sb.value


def poseidon_hash_batch(t, tuples, device=0):
    """Poseidon(t - 1) of every tuple on the GPU, digests only (`hzw_poseidon_hash`)."""
    lib = load_library()
    lib.hzw_poseidon_hash.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p]
    n = len(tuples)
    if n == 0:
        return []
    buf = bytearray(b"".join([x.to_bytes(32, "little") for tup in tuples for x in tup]))
    out = bytearray(32 * n)
    st = lib.hzw_poseidon_hash(device, t, n, _addr(buf), _addr(out))
    if st != 0:
        raise WitnessError("hzw_poseidon_hash failed with code %d" % st)
    return [int.from_bytes(out[32 * i:32 * i + 32], "little") for i in range(n)]

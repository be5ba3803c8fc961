#!/usr/bin/env python
"""Benchmark of the hot path: full witness of rollup-main(nTx, nLevels, maxL1Tx, maxFeeTx).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME]

One step = one pass of the witness generator over one synthetic batch (seeded).
Own arm: `value` = tx/s with the inputs resident in HBM (device time of the kernels, CUDA events
on the launching stream, max over ranks); `e2e` = the same through the public call
`WitnessCalculator.calculateBinWitness` with pinned HOST buffers (input H2D + kernels + witness
D2H inside the timed region).  `--impl reference` times the CPU port of the reference's witness
path (oracle/_ref/libhzw_cpu_port.so; the reference's circom-generated C++ calculator cannot be
built offline) on the host cores, on a bounded strided sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (nTx, nLevels, maxL1Tx, maxFeeTx)
    "rollup-main-2048": (2048, 32, 256, 64),
    "rollup-main-344": (344, 32, 256, 64),
    "rollup-main-24": (24, 32, 8, 4),
    # what one rank of eight holds of the 2048-slot batch (256 slots, 8 fee txs): per-kernel latency floors, not a bench line
    "rollup-main-256": (256, 32, 32, 8),
    "rollup-main-1024": (1024, 32, 128, 32),
    # config 5 of BASELINE.json: 65 536 independent Withdraw(32) witnesses (nTx = instances of the batch)
    "withdraw-65536": (65536, 32, 0, 1),
    "withdraw-4096": (4096, 32, 0, 1),
}
METRIC = "rollup-main(2048,32) witness tx/sec"
CPU_PORT = os.path.join(ROOT, "oracle", "_ref", "libhzw_cpu_port.so")


_REAL_STDOUT = None


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def emit(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def bench_config(cfg, cfg_name):
    """The `config` object of the JSON line: the same in both arms."""
    return {"workload": cfg_name, "nTx": cfg[0], "nLevels": cfg[1], "maxL1Tx": cfg[2], "maxFeeTx": cfg[3],
            "l2_flush": "GPU arm: 256 MiB device memset between timed steps, and every step writes the whole witness "
                        "(far larger than L2); CPU arm: nothing to flush"}


def make_input(cfg, allow_gpu=True):
    """Packed input (NInputs x 32 B) + expected hashGlobalInputs; cached in /tmp for the other arm."""
    import pickle
    nTx, nLevels, maxL1Tx, maxFeeTx = cfg
    # cached beside the repo's sources (git-ignored): building the 2048-slot batch takes the host mirror ~45 s, 3.6 s with its hashes on the GPU
    cdir = os.path.join(ROOT, ".bench_cache")
    try:
        os.makedirs(cdir, exist_ok=True)
    except OSError:
        cdir = tempfile.gettempdir()
    cache = os.path.join(cdir, "hzw_bench_%d_%d_%d_%d.pkl" % cfg)
    if os.path.exists(cache):
        try:
            with open(cache, "rb") as f:
                return pickle.load(f)
        except Exception:
            pass
    from circuits_b200.synth import bench_batch
    t = time.time()
    # with a GPU the trees of the batch are hashed on it, wave by wave (circuits_b200/batchtree.py: same inputs, 3.6 s
    # instead of 45 for the 2048-slot batch and the 16 batches that populate its 4096 accounts); the CPU arm builds on the host
    on_gpu = False
    try:
        import torch
        on_gpu = allow_gpu and torch.cuda.is_available()       # the reference arm touches nothing of the CUDA library
    except Exception:
        pass
    bb = bench_batch(nTx, nLevels, maxL1Tx, maxFeeTx, deferred=on_gpu)
    log("[bench] synthetic batch built in %.1f s (%s)" % (time.time() - t, "hashes on the GPU: %d in %d launches" % (bb.hashes.hashes, bb.hashes.launches) if on_gpu else "host"))
    data = (bb.getInput(), bb.getHashInputs())
    try:
        tmp = cache + ".%d" % os.getpid()
        with open(tmp, "wb") as f:
            pickle.dump(data, f)
        os.replace(tmp, cache)
    except Exception:
        pass
    return data


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.path = os.path.join(tempfile.gettempdir(), "hzw_clocks_%d.csv" % os.getpid())
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if p[4 + k].lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def cpu_baseline(cfg, packed, budget_s=15.0):
    """CPU port on a strided sample of the workload, all host threads."""
    from circuits_b200.witness_calculator import WitnessCalculator
    nTx = cfg[0]
    cores = os.cpu_count() or 1
    wc = WitnessCalculator(*cfg, lib_path=CPU_PORT)
    import numpy as np
    out = np.empty(32 * wc.NVars, dtype=np.uint8)
    # probe with a coarse stride, then size the sample for the budget
    stride = max(nTx // max(cores, 1), 1)
    os.environ["HZW_CPU_SAMPLE_STRIDE"] = str(stride)
    t = time.time()
    wc.calculateBinWitness(packed, False, out=out)
    dt = time.time() - t
    n_s = len(range(0, nTx, stride))
    per_slot_wall = dt / n_s * min(cores, n_s)        # core-seconds per slot
    want = int(budget_s * cores / max(per_slot_wall, 1e-6))
    stride = max(nTx // max(min(want, nTx), 1), 1)
    os.environ["HZW_CPU_SAMPLE_STRIDE"] = str(stride)
    t = time.time()
    wc.calculateBinWitness(packed, False, out=out)
    dt = time.time() - t
    n_s = len(range(0, nTx, stride))
    os.environ.pop("HZW_CPU_SAMPLE_STRIDE", None)
    wc.close()
    import shutil
    # the reference's own calculator needs node (wasm) or circom + a C++ toolchain with ffiasm / gmp (multithreaded C++):
    # probed on the box that runs the bench, reported either way ("kind" stays "port" while neither is there)
    probe = {t: shutil.which(t) for t in ("node", "circom", "snarkjs", "nasm")}
    return {"value": n_s / dt, "unit": "tx/s", "cores": cores, "kind": "port",
            "sample": "every %d-th tx slot / fee tx / SHA block of the %d-slot batch (%d slots, %.1f s wall), "
                      "CPU port of the circuit templates, %d threads" % (stride, nTx, n_s, dt, cores),
            "reference_toolchain_probe": probe}, stride


def run_reference(args, cfg, cfg_name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g
    if not os.path.exists(CPU_PORT):
        g.build()
    from circuits_b200.witness_calculator import WitnessCalculator
    if cfg_name.startswith("withdraw"):
        import numpy as np
        n, nLevels = cfg[0], cfg[1]
        packed, _ = withdraw_batch_inputs(nLevels, n)
        cpu = WitnessCalculator(1, nLevels, 0, 1, circuit="withdraw", lib_path=CPU_PORT)
        mm, nin = min(512, n), cpu.NInputs
        outc = np.empty(32 * cpu.NVars * mm, dtype=np.uint8)
        sample = bytearray(packed[:32 * nin * mm])
        for _ in range(args.warmup):
            cpu.calculateBatch(mm, sample, False, out=outc)
        t = time.time()
        for _ in range(args.steps):
            cpu.calculateBatch(mm, sample, False, out=outc)
        dt = (time.time() - t) / args.steps
        val = mm / dt
        emit({"impl": "reference", "metric": WD_METRIC, "value": val, "unit": "witness/s", "n_gpus": args.gpus, "steps": args.steps,
              "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
              "dtype": "u32 limbs (BN254 Fr, 254-bit modular)", "data": "synthetic",
              "config": {"workload": cfg_name, "instances": n, "nLevels": nLevels, "distinct_leaves": min(WD_DISTINCT, n)},
              "cpu_baseline": {"value": val, "unit": "witness/s", "cores": os.cpu_count() or 1, "kind": "port",
                               "sample": "%d instances of the batch per step, CPU port of the circuit templates, all host threads" % mm},
              "e2e": {"value": val, "unit": "witness/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        return
    inp, _ = make_input(cfg, allow_gpu=False)
    nTx = cfg[0]
    wc = WitnessCalculator(*cfg, lib_path=CPU_PORT)
    packed = wc.packInput(inp)
    wc.close()
    base, stride = cpu_baseline(cfg, packed, budget_s=8.0)
    wc = WitnessCalculator(*cfg, lib_path=CPU_PORT)
    import numpy as np
    out = np.empty(32 * wc.NVars, dtype=np.uint8)
    os.environ["HZW_CPU_SAMPLE_STRIDE"] = str(stride)
    n_s = len(range(0, nTx, stride))
    for _ in range(args.warmup):
        wc.calculateBinWitness(packed, False, out=out)
    t = time.time()
    for _ in range(args.steps):
        wc.calculateBinWitness(packed, False, out=out)
    dt = (time.time() - t) / args.steps
    os.environ.pop("HZW_CPU_SAMPLE_STRIDE", None)
    val = n_s / dt
    base["value"] = val
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "tx/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u32 limbs (BN254 Fr, 254-bit modular)",
        "data": "synthetic", "config": bench_config(cfg, cfg_name),
        "cpu_baseline": base,
        "e2e": {"value": val, "unit": "tx/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0})


def poseidon_metric(device):
    """Second metric of BASELINE.json: Poseidon permutations per second (t = 3, every signal of the
    component stored, as inside the SMT levels), 65 536 lane groups = 131 072 permutations per launch."""
    import random
    import ctypes
    from circuits_b200 import witness_calculator as wcm
    from circuits_b200.field import P
    lib = wcm.load_library()
    lib.hzw_poseidon_bench.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.POINTER(ctypes.c_float)]
    n = 65536
    rng = random.Random(20261017)
    buf = bytearray(b"".join(rng.randrange(P).to_bytes(32, "little") for _ in range(4 * n)))
    ms = ctypes.c_float()
    res = {}
    for name, reps in (("transposed", 5), ("lane_groups", -5)):
        st = lib.hzw_poseidon_bench(device, n, reps, wcm._addr(buf), None, ctypes.byref(ms))
        if st != 0:
            return None
        res[name] = ms.value
    size = 634
    best = min(res.values())
    # the other widths of the path (t = 4 leaf hashes, 5 state hashes, 6 the verifier's hash, 7 the message hash)
    per_width = {"3": {"perm_per_s": 2 * n / (best * 1e-3), "ms_per_launch": best, "perms_per_launch": 2 * n}}
    try:
        for t in (4, 5, 6, 7):
            tuples = [[rng.randrange(P) for _ in range(t - 1)] for _ in range(1024)] * 64
            _, ms_t, sb = wcm.poseidon_width(t, tuples, reps=3, device=device)
            per_width[str(t)] = {"perm_per_s": len(tuples) / (ms_t * 1e-3), "ms_per_launch": ms_t, "perms_per_launch": len(tuples),
                                 "hbm_gbs": sb / (ms_t * 1e-3) / 1e9}
    except Exception as e:      # a measurement helper must not take the bench line down
        log("[bench] poseidon widths skipped: %r" % (e,))
    return {"value": 2 * n / (best * 1e-3), "unit": "perm/s", "t": 3, "perms_per_launch": 2 * n,
            "ms_per_launch": best, "signal_bytes_per_launch": 2 * n * size * 32,
            "hbm_gbs": 2 * n * size * 32 / (best * 1e-3) / 1e9,
            "kernels_ms": res, "per_width": per_width}


def poseidon_perm_products(t):
    """32x32 limb products of one Poseidon permutation of width t as the kernels evaluate it."""
    rp = {3: 57, 4: 56, 5: 60, 6: 60, 7: 63}[t]
    return (8 * t + rp) * 3 * 153 + (8 + rp) * t * (64 * t + 100)


def poseidon_products(cfg):
    nTx, nLevels, _, F = cfg
    per_proc = 2 * (nLevels + 1) * poseidon_perm_products(3) + 2 * poseidon_perm_products(4)
    per_tx = 2 * per_proc + 4 * poseidon_perm_products(5) + poseidon_perm_products(6) + poseidon_perm_products(7)
    per_fee = per_proc + 2 * poseidon_perm_products(5)
    return nTx * per_tx + F * per_fee


def committed_traffic(kernel, cfg_name):
    """dram bytes per launch of the dominant kernel, from the ncu --set full capture summarised under profiles/."""
    try:
        with open(os.path.join(ROOT, "profiles", "r2_dram_traffic.json")) as f:
            t = json.load(f)
        return t.get(cfg_name, {}).get(kernel)
    except Exception:
        return None


def run_gpu(args, cfg, cfg_name):
    import torch
    import __graft_entry__ as g
    from circuits_b200 import witness_calculator as wcm
    if not os.path.exists(wcm.DEFAULT_LIB):
        g.build()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        log("[bench] WORLD_SIZE=%d but --gpus %d; using WORLD_SIZE" % (world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nTx = cfg[0]
    inp, expect_hash = make_input(cfg)
    sharded = world > 1
    if sharded:
        wc = wcm.ShardedWitnessCalculator(*cfg, rank=rank, world=world, device=local)
        out_bytes = 32 * wc.shardNVars
    else:
        wc = wcm.WitnessCalculator(*cfg, device=local)
        out_bytes = 32 * wc.NVars
    info = wc.planInfo()
    pin_in = torch.empty(32 * wc.NInputs, dtype=torch.uint8, pin_memory=True)
    pin_in.numpy()[:] = memoryview(wc.packInput(inp))
    pin_out = torch.empty(out_bytes, dtype=torch.uint8, pin_memory=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def resident_step(first=False):
        if sharded:
            return wc.calculateShard(pin_in if first else None, False, resident=True)[1]
        return wc.calculateResident(pin_in if first else None, False)

    def e2e_step():
        if sharded:
            wc.calculateShard(pin_in, False, out=pin_out)
        else:
            wc.calculateBinWitness(pin_in, False, out=pin_out)

    # -- device-resident runs: inputs already in HBM, witness left in HBM ------------------------
    # clocks / throttle reasons are sampled from the warm-up to the end of the end-to-end loop: nvidia-smi needs a
    # few hundred ms to start and the device-resident timed region alone is only steps x ~13 ms long
    sampler = ClockSampler(local) if rank == 0 else None
    resident_step(first=True)                      # uploads the inputs once
    for _ in range(max(args.warmup - 1, 0)):
        resident_step()
    barrier()
    step_ms, kern = [], {}
    for _ in range(args.steps):
        flush.zero_()                               # evict L2 between timed steps
        barrier()
        ms = resident_step()
        step_ms.append(ms)
        for name, kms in wc.lastKernelMs():
            kern.setdefault(name, []).append(kms)
    barrier()
    t_dev = torch.tensor([sum(step_ms) / len(step_ms)], device="cuda")
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_dev.item())
    # -- end to end through the public call, pinned host buffers ---------------------------------
    # One GPU: WitnessCalculator.calculateBinWitness.  N GPUs: the same call on ONE context that spans the N GPUs of
    # the box (hzw_params.n_gpus = N; rank 0's process drives them, the exchange runs inside the library and one
    # whole witness lands in one pinned buffer) -- the other ranks wait on a host-side (gloo) barrier meanwhile, so
    # that no NCCL kernel spins on their GPUs.
    if sharded:
        del pin_out
        host_group = dist.new_group(backend="gloo")
        e2e_t, stats, got_hash = [], None, None
        if rank == 0:
            wm = wcm.WitnessCalculator(*cfg, device=0, n_gpus=world)
            pin_full = torch.empty(32 * wm.NVars, dtype=torch.uint8, pin_memory=True)
            wm.calculateBinWitness(pin_in, False, out=pin_full)
            for _ in range(args.steps):
                torch.cuda.synchronize()
                t = time.perf_counter()
                wm.calculateBinWitness(pin_in, False, out=pin_full)
                e2e_t.append(time.perf_counter() - t)
            stats = wm.lastStats()
            got_hash = int.from_bytes(bytes(pin_full[32:64].numpy()), "little")
            wm.close()
            del pin_full
        dist.barrier(group=host_group)
        clocks = sampler.stop() if sampler else None
        if rank != 0:
            return
        e2e_s = sum(e2e_t) / len(e2e_t)
        io = [float(stats["h2d_bytes"]), float(stats["d2h_bytes"])]
        launches_per_step = int(stats["gpu_launches"])
    else:
        e2e_step()
        barrier()
        e2e_t = []
        for _ in range(args.steps):
            barrier()
            t = time.perf_counter()
            e2e_step()
            e2e_t.append(time.perf_counter() - t)
        barrier()
        clocks = sampler.stop() if sampler else None
        stats = wc.lastStats()
        e2e_s = sum(e2e_t) / len(e2e_t)
        io = [float(stats["h2d_bytes"]), float(stats["d2h_bytes"])]
        launches_per_step = int(stats["gpu_launches"])
        got_hash = int.from_bytes(bytes(pin_out[32:64].numpy()), "little")
    # hashGlobalInputs is witness entry 1
    if got_hash != expect_hash:
        raise SystemExit("hashGlobalInputs mismatch: the witness is wrong, refusing to report a number")
    if rank != 0:
        return
    # -- roofline (rank 0's launches) ---------------------------------------------------------------
    # Per-kernel durations: at one GPU from hzw_profile_step (one more batch with every launch serialised on one
    # stream, an event in front of each: the time a kernel takes ALONE, measured live); on a sharded context from the
    # intervals of the timed steps (five streams run together there, so an interval also holds its neighbours).
    kmean = {k: sum(v) / len(v) for k, v in kern.items()}
    F, L1 = cfg[3], cfg[1] + 1
    n_slots = (nTx + world - 1) // world
    n_fee = (F + world - 1) // world
    n_proc = 2 * n_slots + n_fee
    SZL = 2 + 4 + 2 * 634 + 3 + 5                  # entries of one SMTProcessorLevel (hzw_dev.cuh HZ_SZL)
    alone, lev_tot, lev_img = None, n_proc * L1, 0
    if not sharded:
        wc.profileStep()
        runs = [wc.profileStep() for _ in range(3)]
        alone = {}
        for r in runs:
            for name, ms in r:
                alone[name] = alone.get(name, 0.0) + ms / len(runs)
        lev_tot, lev_img = wc.imageLevels()
    kms = alone if alone is not None else kmean
    top = max(kms, key=kms.get)
    # two kernels share the top within a few percent from run to run (k_smt_levels, k_front): the one with a counted
    # integer workload is reported when it is within 10 % of the longest, so that the line does not flip between rounds
    for name in ("k_smt_levels",):
        if name in kms and kms[name] >= 0.9 * kms[top]:
            top = name
    fee_proc = info["feeTx"] - 2 * 980 - 30       # entries of a FeeTx spent in its SMT processor (approx.)
    front_entries = info["decodeTx"] + info["rollupTx"] - info["getAx"] - info["sigVerifier"] - info["processors"]
    sha_blocks = (info["sha_blocks"] + world - 1) // world
    # ALGORITHMIC bytes per launch: the witness entries a kernel has to write + the main inputs it reads, 32 B each
    alg_bytes = {
        "k_front": 32 * n_slots * (front_entries + 362 + 2 * F),
        "k_smtjob": 32 * n_slots * (362 + 2 * F),      # reads only: a dry run for the SMT job records
        "k_smt_levels": 32 * ((lev_tot - lev_img) * SZL + n_proc * L1),
        "k_smt_zero": 32 * lev_img * 2 * 634,
        "k_smt": 32 * n_proc * (info["processors"] // 2 - L1 * SZL) if alone is not None else
                 32 * (n_slots * (info["processors"] + 2 * L1 + 14) + n_fee * (fee_proc + L1 + 7)),
        "k_ed_fill": 32 * n_slots * (252 * 9 + 83 * 3 + 83 * 17),
        "k_eddsa": 32 * n_slots * (info["getAx"] + info["sigVerifier"] + 8),
        "k_hstate": 32 * (4 * n_slots + 2 * n_fee) * 980,
        "k_leafhash": 32 * 2 * n_proc * 776,
        "k_sighash": 32 * n_slots * 1351,
        "k_sha_blocks": 32 * sha_blocks * info["sha_block"],
    }
    # ALGORITHMIC limb products (32 x 32 -> 64) per launch of the Poseidon kernels
    alg_prod = {
        "k_smt_levels": (lev_tot - lev_img) * 2 * poseidon_perm_products(3),
        "k_hstate": (4 * n_slots + 2 * n_fee) * poseidon_perm_products(5),
        "k_leafhash": 2 * n_proc * poseidon_perm_products(4),
        "k_sighash": n_slots * poseidon_perm_products(7),
    }
    peaks, src = measured_peaks()
    try:
        ipeak = wcm.int_peak(local)
    except Exception as e:      # a measurement helper must not take the bench line down
        log("[bench] int peak skipped: %r" % (e,))
        ipeak = None
    per_kernel = {}
    for name, ms in kms.items():
        e = {"ms": ms}
        if alg_bytes.get(name) and ms > 0:
            e["algorithmic_bytes"] = alg_bytes[name]
            e["hbm_gbs"] = alg_bytes[name] / (ms * 1e-3) / 1e9
            e["hbm_frac"] = e["hbm_gbs"] / peaks["hbm_gbs"]
        if alg_prod.get(name) and ms > 0 and ipeak:
            e["limb_products"] = alg_prod[name]
            e["int_frac"] = alg_prod[name] / (ms * 1e-3) / ipeak
        per_kernel[name] = e
    tk = per_kernel[top]
    int_bound = "int_frac" in tk
    roofline = {
        "bound": "int" if int_bound else "hbm", "kernel": top,
        "achieved": (tk["limb_products"] / (tk["ms"] * 1e-3)) if int_bound else tk.get("hbm_gbs", 0.0),
        "peak": ipeak if int_bound else peaks["hbm_gbs"],
        "unit": "32x32->64 multiply-adds/s over all lanes" if int_bound else "GB/s",
        "frac": tk["int_frac"] if int_bound else tk.get("hbm_frac", 0.0),
        "traffic": committed_traffic(top, cfg_name),
        "peak_source": "measured live (hzw_int_peak: independent IMAD.WIDE.U32 chains, full occupancy)" if int_bound
                       else src + " (copy bandwidth)",
        "hbm": {"achieved": tk.get("hbm_gbs"), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": tk.get("hbm_frac"),
                "algorithmic_bytes": tk.get("algorithmic_bytes"), "peak_source": src + " (copy bandwidth)"},
        "kernel_ms_source": "hzw_profile_step: each launch alone on one stream, CUDA events, mean of 3 batches" if alone is not None
                            else "intervals of the timed steps (CUDA events on the launching stream; streams run together)",
        "kernels": per_kernel, "timed_step_intervals_ms": kmean,
        "image_levels": {"processor_levels": lev_tot, "copied_from_the_(0,0)_image": lev_img},
        "step": {"witness_bytes": 32 * info["nvars"], "hbm_gbs": 32 * info["nvars"] / world / (ms_per_step * 1e-3) / 1e9,
                 "hbm_frac": 32 * info["nvars"] / world / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"],
                 "sum_of_kernels_alone_ms": sum(kms.values()) if alone is not None else None},
        "note": "254-bit modular arithmetic: the dominant kernels are bound by dependent integer issue (IMAD.WIDE carry "
                "chains), so the primary bound is the integer-multiply pipe and the HBM fraction is the secondary figure "
                "(DESIGN.md section 7). Algorithmic work of k_smt_levels / k_smt_zero follows the (0, 0) levels of THIS batch.",
    }
    base = None
    if world == 1 and not args.no_cpu_baseline:
        if not os.path.exists(CPU_PORT):
            g.build()
        base, _ = cpu_baseline(cfg, bytearray(pin_in.numpy().tobytes()))
    line = {
        "metric": METRIC, "value": nTx / (ms_per_step * 1e-3), "unit": "tx/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u32 limbs (BN254 Fr, 254-bit modular)", "data": "synthetic",
        "config": bench_config(cfg, cfg_name),
        "workload_detail": {"witness_signals": info["nvars"], "witness_bytes": 32 * info["nvars"],
                            "sharding": "tx slots / fee txs / SHA blocks split over %d rank(s); one all-gather of %s B records"
                                        % (world, getattr(wc, "recBytes", 0)) if sharded else "single GPU"},
        "e2e": {"value": nTx / e2e_s, "unit": "tx/s", "h2d_bytes_per_step": int(io[0]),
                "d2h_bytes_per_step": int(io[1]), "ms_per_step": e2e_s * 1e3,
                "call": "WitnessCalculator.calculateBinWitness -> hzw_calculate, pinned host buffers, ONE whole witness out"
                        + (" (one context with n_gpus = %d: exchange inside the library, every GPU copies its ranges "
                           "to their offsets in the one buffer)" % world if sharded else "")},
        "gpu_launches": launches_per_step * args.steps if sharded else launches_per_step * args.steps * world,
        "clocks": clocks, "roofline": roofline,
    }
    if world == 1:
        pm = poseidon_metric(local)
        if pm is not None:
            line["poseidon"] = pm
        if ipeak:
            peak = ipeak
            prod = poseidon_products(cfg)
            skipped = lev_img * 2 * poseidon_perm_products(3)
            # EdDSA per signature (hints.cuh): 253 + 253 + 83 point operations of 14 products each on the ladder walk, the
            # same formulas once more when the step entries are written, 3 per element of the batched inversion (669),
            # the square root of getAx (~1.5 k), Poseidon(5) counted above with the Poseidon kernels
            ed_mulmods = (253 + 253 + 83) * 14 * 2 + 3 * 669 + 1500
            ed = n_slots * world * ed_mulmods * 153
            line["int_roofline"] = {
                "unit": "32x32->64 multiply-adds/s over all lanes", "peak": peak,
                "peak_source": "measured live (hzw_int_peak: independent IMAD.WIDE.U32 chains, full occupancy)",
                "achieved": (prod - skipped + ed) / (ms_per_step * 1e-3),
                "frac": (prod - skipped + ed) / (ms_per_step * 1e-3) / peak,
                "executed_products_per_step": prod - skipped + ed,
                "poseidon_products_of_the_circuit": prod, "poseidon_products_skipped_as_images": skipped,
                "eddsa_products_estimate": ed,
                "frac_counting_the_skipped_levels": (prod + ed) / (ms_per_step * 1e-3) / peak,
                "note": "whole step: Poseidon (S-boxes: 153 limb products per modular multiplication; MDS rows: 64 t + 100 "
                        "per row) without the (0, 0) levels this batch copied from the image, plus the EdDSA estimate; "
                        "SHA-256 and bit plumbing not counted"}
            if pm is not None:
                pp = pm["perms_per_launch"] * poseidon_perm_products(3)
                line["poseidon"]["int_frac"] = pp / (pm["ms_per_launch"] * 1e-3) / peak
    if base is not None:
        line["cpu_baseline"] = base
    emit(line)


WD_DISTINCT = 4096
WD_METRIC = "withdraw(32) witnesses/sec"


def withdraw_batch_inputs(nLevels, n):
    """n packed Withdraw inputs: WD_DISTINCT distinct leaves of ONE exit tree with their inclusion proofs (the host
    mirror of withdrawUtils builds them with big-integer Poseidon: ~1 minute, cached), repeated to n instances --
    every instance is evaluated independently, the work per instance does not depend on its values."""
    import pickle
    cdir = os.path.join(ROOT, ".bench_cache")
    try:
        os.makedirs(cdir, exist_ok=True)
    except OSError:
        cdir = tempfile.gettempdir()
    cache = os.path.join(cdir, "hzw_withdraw_%d_%d.pkl" % (nLevels, WD_DISTINCT))
    data = None
    if os.path.exists(cache):
        try:
            with open(cache, "rb") as f:
                data = pickle.load(f)
        except Exception:
            data = None
    if data is None:
        from circuits_b200.synth import withdraw_inputs
        t = time.time()
        cases = withdraw_inputs(nLevels, WD_DISTINCT)
        names = ["rootExit", "ethAddr", "tokenID", "balance", "idx", "sign", "ay"]
        packed = bytearray()
        for inp, _ in cases:
            vals = [inp[k] for k in names] + list(inp["siblingsState"])
            packed += b"".join(int(v).to_bytes(32, "little") for v in vals)
        data = (bytes(packed), [h for _, h in cases])
        log("[bench] %d withdraw inputs built in %.1f s" % (WD_DISTINCT, time.time() - t))
        try:
            tmp = cache + ".%d" % os.getpid()
            with open(tmp, "wb") as f:
                pickle.dump(data, f)
            os.replace(tmp, cache)
        except Exception:
            pass
    packed, hashes = data
    reps = (n + WD_DISTINCT - 1) // WD_DISTINCT
    nin = 8 + nLevels
    return (packed * reps)[:32 * nin * n], (hashes * reps)[:n]


def run_withdraw(args, cfg, cfg_name):
    """Config 5: `n` independent Withdraw(nLevels) witnesses, the instances split over the ranks (no exchange).
    value = witnesses/s with the witnesses left in HBM (only the public output of each comes back and is checked);
    e2e = the same through calculateBatch with every witness copied to a pinned host buffer, on a bounded sub-batch
    (the 65 536 witnesses are 197 GB)."""
    import numpy as np
    import torch
    import __graft_entry__ as g
    from circuits_b200 import witness_calculator as wcm
    if not os.path.exists(wcm.DEFAULT_LIB):
        g.build()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, nLevels = cfg[0], cfg[1]
    packed, hashes = withdraw_batch_inputs(nLevels, n)
    per = (n + world - 1) // world
    k0, k1 = min(rank * per, n), min(rank * per + per, n)
    wc = wcm.WitnessCalculator(1, nLevels, 0, 1, circuit="withdraw", device=local)
    nin = wc.NInputs
    pin_in = torch.empty(32 * nin * (k1 - k0), dtype=torch.uint8, pin_memory=True)
    pin_in.numpy()[:] = np.frombuffer(packed, dtype=np.uint8)[32 * nin * k0:32 * nin * k1]
    pub = torch.empty(32 * (k1 - k0), dtype=torch.uint8, pin_memory=True)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 1)):
        wc.calculateBatchResident(k1 - k0, pin_in, False, public_out=pub)
    barrier()
    step_ms = []
    for _ in range(args.steps):
        barrier()
        _, ms = wc.calculateBatchResident(k1 - k0, pin_in, False, public_out=pub)
        step_ms.append(ms)
    barrier()
    got = pub.numpy().reshape(-1, 32)
    for j in range(0, k1 - k0, max((k1 - k0) // 257, 1)):
        if int.from_bytes(bytes(got[j]), "little") != hashes[k0 + j]:
            raise SystemExit("hashGlobalInputs of instance %d is wrong, refusing to report a number" % (k0 + j))
    launches = wc.lastStats()["gpu_launches"]
    t_dev = torch.tensor([sum(step_ms) / len(step_ms)], device="cuda")
    if dist is not None:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_dev.item())
    # end to end on a bounded sub-batch per rank
    m = min(k1 - k0, 1024)
    pin_out = torch.empty(32 * wc.NVars * m, dtype=torch.uint8, pin_memory=True)
    wc.calculateBatch(m, pin_in[:32 * nin * m], False, out=pin_out)
    e2e_t = []
    for _ in range(args.steps):
        barrier()
        t = time.perf_counter()
        wc.calculateBatch(m, pin_in[:32 * nin * m], False, out=pin_out)
        e2e_t.append(time.perf_counter() - t)
    barrier()
    clocks = sampler.stop() if sampler else None
    t_e2e = torch.tensor([sum(e2e_t) / len(e2e_t)], device="cuda")
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    if rank != 0:
        return
    e2e_s = float(t_e2e.item())
    peaks, src = measured_peaks()
    try:
        ipeak = wcm.int_peak(local)
    except Exception:
        ipeak = None
    L1 = nLevels + 1
    prod = per * (L1 * poseidon_perm_products(3) + 2 * poseidon_perm_products(4) + poseidon_perm_products(5))
    alg_bytes = per * 32 * (wc.NVars + nin)
    roofline = {"bound": "hbm", "kernel": "k_withdraw", "achieved": alg_bytes / (ms_per_step * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": alg_bytes / (ms_per_step * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                "peak_source": src + " (copy bandwidth)", "algorithmic_bytes": alg_bytes,
                "int": {"limb_products": prod, "frac": prod / (ms_per_step * 1e-3) / ipeak if ipeak else None, "peak": ipeak},
                "note": "one kernel per chunk of 4096 instances; algorithmic bytes = every witness entry written once + the "
                        "inputs; 62 % of the entries of an instance are the bit signals of its two SHA-256 blocks"}
    line = {"metric": WD_METRIC, "value": n / (ms_per_step * 1e-3), "unit": "witness/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u32 limbs (BN254 Fr, 254-bit modular)", "data": "synthetic",
            "config": {"workload": cfg_name, "instances": n, "nLevels": nLevels,
                       "distinct_leaves": min(WD_DISTINCT, n),
                       "l2_flush": "every step writes %d GB of witnesses per GPU, far larger than L2" % (32 * wc.NVars * per >> 30)},
            "e2e": {"value": m * world / e2e_s, "unit": "witness/s", "h2d_bytes_per_step": 32 * nin * m * world,
                    "d2h_bytes_per_step": 32 * wc.NVars * m * world, "ms_per_step": e2e_s * 1e3,
                    "call": "WitnessCalculator.calculateBatch on a sub-batch of %d instances per GPU, witnesses to pinned host memory" % m},
            "gpu_launches": int(launches) * args.steps * world, "clocks": clocks, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        cpu = wcm.WitnessCalculator(1, nLevels, 0, 1, circuit="withdraw", lib_path=CPU_PORT)
        mm = 512
        outc = np.empty(32 * cpu.NVars * mm, dtype=np.uint8)
        t = time.time()
        cpu.calculateBatch(mm, bytearray(packed[:32 * nin * mm]), False, out=outc)
        dt = time.time() - t
        line["cpu_baseline"] = {"value": mm / dt, "unit": "witness/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "%d instances of the batch, CPU port of the circuit templates, all host threads, %.1f s" % (mm, dt)}
    emit(line)


def main():
    # the driver reads ONE JSON line from stdout: libraries that print there (NCCL's version banner) go to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="rollup-main-2048", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, cfg, args.config)
    elif args.config.startswith("withdraw"):
        run_withdraw(args, cfg, args.config)
    else:
        run_gpu(args, cfg, args.config)
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                dist.destroy_process_group()
        except Exception:
            pass


if __name__ == "__main__":
    main()

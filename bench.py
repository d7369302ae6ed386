#!/usr/bin/env python
"""bench.py -- bootstrapped NAND gates/s (SECURITY_128_BIT) on N B200s.

A step = one pass of the hot path (gates::batch_nand: prep -> blind rotation ->
sample extract -> key switch) over one batch of synthetic ciphertexts.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--count C] [--params 128]
  python bench.py --impl reference ...     # CPU arm: the oracle port on host cores

N>1 is launched by the driver with torch.distributed.run (one rank per GPU).
Prints ONE JSON line on rank 0.  `value`: inputs resident in HBM, device-timed.
`e2e`: the same metric through the host-buffer C-ABI call (tfhe_batch_gate) on ordinary PAGEABLE
buffers -- what a drop-in caller's Vec<Ciphertext> is -- with H2D/D2H inside the timed region
(`e2e_pinned`: the same with pinned buffers).  `configs`: the other BASELINE configurations (C1 at
its literal 1024 gates, C2 ends, C3, C4).  `c5`: 1 048 576 mixed gates sharded over the ranks (strong
scaling) with the warm key broadcast.  See DESIGN.md section "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "bootstrapped_nand_gates_per_sec_128bit"
UNIT = "gates/s"
N = 1024


def flop_per_pbs(n: int, l: int) -> float:
    """n*[(2l+2)*(5*(N/2)*log2(N/2) + 6*(N/2)) + 2l*2*(N/2)*8]."""
    h = N // 2
    return n * ((2 * l + 2) * (5 * h * 9 + 6 * h) + 2 * l * 2 * h * 8)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--count", type=int, default=65536, help="gates per GPU per step")
    ap.add_argument("--params", default="128")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="gates in the CPU sample (0: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(params: str, count: int) -> str:
    return (f"gates::batch_nand, SECURITY_{params.upper()}_BIT, {count} gates per GPU per step "
            f"(BASELINE configs[0] shape at a throughput batch size)")


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2])); pw.append(float(parts[3]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        top = max(sm)
        load = [x for x in sm if x >= 0.5 * top]
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(pw), "samples": len(sm)}


# ----------------------------------------------------------------------------- CPU arm
def cpu_leg(params: str, sample: int, threads: int = 0, engine=None):
    """The oracle port (C restatement of rs-tfhe's Rayon path; rs-tfhe itself cannot be
    built here: no cargo/rustc) timed on the host cores over a bounded sample of the same
    workload.  If `engine` is given the sample is also pushed through the GPU path with
    the oracle's (real) key and compared word for word -- the oracle acting as checker."""
    import oracle as O

    if threads <= 0:
        threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else O.max_threads()
    cores = threads
    K = O.Keys(params, seed=0x5EED0001)
    if sample <= 0:
        sample = cores * 4
        r = np.random.default_rng(5)
        probe = r.integers(0, 2**32, (cores * 2, 2, K.params.n + 1), dtype=np.uint32)
        t = time.perf_counter()
        K.batch_gate(0, probe, threads=cores)
        per = (time.perf_counter() - t) / (cores * 2)
        sample = int(max(cores * 4, min(cores * 512, 12.0 / per)))
        sample -= sample % cores
    r = np.random.default_rng(6)
    pairs = r.integers(0, 2**32, (sample, 2, K.params.n + 1), dtype=np.uint32)
    t = time.perf_counter()
    ref = K.batch_gate(0, pairs, threads=cores)
    dt = time.perf_counter() - t
    out = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
           "sample": f"{sample} random-ciphertext NAND gates of the same workload, "
                     f"{cores} OpenMP threads (one ciphertext per task, like par_map)",
           "seconds": dt, "ms_per_gate_per_core": dt * cores / sample * 1e3}
    if engine is not None:
        import rs_tfhe_b200 as T
        ck = T.CloudKey(T.PARAMS_BY_NAME[params], K.offset, K.tv_a, K.tv_b, K.ksk, K.bsk)
        engine.load_cloud_key(ck)
        got = engine.batch_gate("NAND", pairs)
        out["parity_mismatch_words"] = int((got != ref).sum())
        out["parity_checked_gates"] = sample
    return out


def parity_sample(params: str, engine, gates: int, threads: int, seed: int):
    """Real-key parity on THIS rank: `gates` random mixed gates through the engine's current device
    (after the caller loaded / received the real key) against the oracle, word for word."""
    import oracle as O
    K = O.Keys(params, seed=0x5EED0001)
    r = np.random.default_rng(seed)
    pairs = r.integers(0, 2**32, (gates, 2, K.params.n + 1), dtype=np.uint32)
    ops = r.integers(0, 6, gates).astype(np.uint8)
    ref = K.batch_gate(ops, pairs, threads=max(1, threads))
    return K, pairs, ops, ref


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as O
    # all host threads, also under torchrun (which exports OMP_NUM_THREADS=1)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    K = O.Keys(args.params, seed=0x5EED0001)
    sample = args.cpu_sample or cores * 8
    r = np.random.default_rng(7)
    pairs = r.integers(0, 2**32, (sample, 2, K.params.n + 1), dtype=np.uint32)
    for _ in range(args.warmup):
        K.batch_gate(0, pairs[:cores], threads=cores)
    t = time.perf_counter()
    for _ in range(args.steps):
        K.batch_gate(0, pairs, threads=cores)
    dt = time.perf_counter() - t
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args.params, args.count), "params": args.params,
                   "count_per_gpu": args.count,
                   "note": "CPU arm: C restatement of rs-tfhe's Rayon path (oracle/, kind=port; the "
                           "Rust crate cannot be built in this image), all host threads; each step "
                           "is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} NAND gates per step x {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)


# ----------------------------------------------------------------------------- GPU arm
def synthetic_cloud_key(T, P, seed: int):
    """Random key material in the reference's layout.  The data path is value-independent
    (no branch depends on key values), so throughput equals that of a real key; real keys
    are used by tests/, smoke() and the cpu_baseline parity check."""
    r = np.random.default_rng(seed)
    ksk = r.integers(0, 2**32, (P.ksk_rows, P.n + 1), dtype=np.uint32)
    # Fourier image (klemsa.rs:88-117: twist, 512-point FFT, x2) of uniform random torus rows,
    # so the arithmetic stays in the exact-integer regime exactly as with a real key
    x = r.integers(-2**31, 2**31, (P.n * 2 * P.l * 2, N)).astype(np.float64)
    z = (x[:, :N // 2] + 1j * x[:, N // 2:]) * np.exp(1j * np.pi * np.arange(N // 2) / N)
    F = np.fft.fft(z, axis=1) * 2.0
    bsk = np.concatenate([F.real, F.imag], axis=1).reshape(P.n, 2 * P.l, 2, N)
    tv_a = np.zeros(N, dtype=np.uint32)
    tv_b = np.full(N, 0x20000000, dtype=np.uint32)
    offset = sum((1 << (P.bgbit - 1)) << (32 - (i + 1) * P.bgbit) for i in range(P.l)) & 0xFFFFFFFF
    return T.CloudKey(P, offset, tv_a, tv_b, ksk, bsk)


_JSON_FD = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version
    banner to fd 1), so fd 1 is pointed at stderr for the whole run and the JSON line goes to a
    private duplicate of the original stdout."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def time_host_call(fn, reps: int):
    """best and mean wall time (s) of a host-buffer call"""
    ts = []
    for _ in range(reps):
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return min(ts), sum(ts) / len(ts)


def other_configs(T, torch, dev, local, warm_engine):
    """BASELINE configs beside the bench workload, each a few hundred ms: C1 at its literal 1024 gates,
    C2 (80/110-bit, 1k and 64k gates), C3 (UINT4 programmable bootstrap, 16k) and C4 (the nibble adder's
    dependent PBS chain).  Host-buffer numbers use pageable numpy arrays."""
    out = {}
    r = np.random.default_rng(42)

    def gates_cfg(eng, P, count, reps):
        pairs = r.integers(0, 2**32, (count, 2, P.n + 1), dtype=np.uint32)
        eng.batch_gate("NAND", pairs)                       # warm (allocations)
        best, mean = time_host_call(lambda: eng.batch_gate("NAND", pairs), reps)
        br, ks = eng.last_kernel_ms()
        return {"count": count, "e2e_gates_per_s": count / best, "e2e_ms": best * 1e3, "e2e_ms_mean": mean * 1e3,
                "kernels_ms": {"blind_rotate": br, "key_switch": ks}, "kernels_gates_per_s": count / ((br + ks) * 1e-3)}

    # C1: examples/batch_gates.rs:53-78 -- 1024 hom_nand gates, 128-bit
    P128 = T.PARAMS_BY_NAME["128"]
    out["C1_1024_nand_128bit"] = gates_cfg(warm_engine, P128, 1024, 8)
    # C2: 80-bit and 110-bit sweep ends
    for name in ("80", "110"):
        P = T.PARAMS_BY_NAME[name]
        e = T.CudaBootstrap(P, local)
        e.load_cloud_key(synthetic_cloud_key(T, P, 77))
        out[f"C2_{name}bit"] = {"1k": gates_cfg(e, P, 1024, 5), "64k": gates_cfg(e, P, 65536, 2)}
        e.close()
    # C3: LutBootstrap::bootstrap_func, SECURITY_UINT4, messageModulus 16, batch 16384
    P = T.PARAMS_BY_NAME["uint4"]
    e = T.CudaBootstrap(P, local)
    e.load_cloud_key(synthetic_cloud_key(T, P, 78))
    cts = r.integers(0, 2**32, (16384, P.n + 1), dtype=np.uint32)
    table = [(x * x) % 16 for x in range(16)]
    e.batch_bootstrap_func(table, 16, cts)
    best, mean = time_host_call(lambda: e.batch_bootstrap_func(table, 16, cts), 4)
    br, ks = e.last_kernel_ms()
    out["C3_uint4_lut_16k"] = {"count": 16384, "e2e_pbs_per_s": 16384 / best, "e2e_ms": best * 1e3,
                               "kernels_ms": {"blind_rotate": br, "key_switch": ks},
                               "kernels_pbs_per_s": 16384 / ((br + ks) * 1e-3)}
    e.close()
    # C4: examples/lut_add_two_numbers.rs:80-157 -- 3 PBS in 2 dependent levels, 128-bit gate params, modulus 32
    gen = T.Generator(32, warm_engine)
    lut_low = gen.generate_lookup_table(lambda x: x % 16)
    lut_carry = gen.generate_lookup_table(lambda x: 1 if x >= 16 else 0)
    nib = r.integers(0, 2**32, (4, P128.n + 1), dtype=np.uint32)

    def chain():
        lo = (nib[0] + nib[2]).astype(np.uint32)
        both = warm_engine.batch_bootstrap_lut([lut_low.lut_id, lut_carry.lut_id], np.stack([lo, lo]))
        hi = (nib[1] + nib[3] + both[1]).astype(np.uint32)
        return warm_engine.batch_bootstrap_lut(lut_low.lut_id, hi)

    chain()
    best, mean = time_host_call(chain, 20)
    one = r.integers(0, 2**32, (1, P128.n + 1), dtype=np.uint32)
    warm_engine.batch_bootstrap_lut(lut_low.lut_id, one)
    b1, _ = time_host_call(lambda: warm_engine.batch_bootstrap_lut(lut_low.lut_id, one), 20)
    out["C4_nibble_add_chain"] = {"levels": 2, "pbs": 3, "latency_ms": best * 1e3, "latency_ms_mean": mean * 1e3,
                                  "single_pbs_ms": b1 * 1e3, "us_per_pbs_in_chain": best * 1e6 / 3}
    lut_low.release(); lut_carry.release()
    return out


def main():
    args = parse_args()
    _claim_stdout()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import ctypes as C

    import torch
    import torch.distributed as dist

    import rs_tfhe_b200 as T
    from rs_tfhe_b200.dist import broadcast_cloud_key, shard_range

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run for --gpus > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    P = T.PARAMS_BY_NAME[args.params]
    w = P.n + 1
    count = args.count
    eng = T.CudaBootstrap(P, local)
    stream = torch.cuda.Stream(device=dev)
    eng.set_stream(stream.cuda_stream)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- parity on EVERY rank with the real (oracle) key: rank 0 uploads it, the others receive the
    # re-laid-out blob by NCCL broadcast and rebuild their kernel-specific key orders (commit path)
    K = None
    parity = {"mismatch_words": None, "gates": 0}
    first_bcast_ms = None
    if not args.no_cpu_baseline:
        import oracle as O
        gates = 64 if world > 1 else 0       # at N=1 the cpu_baseline leg below checks its whole sample
        if world > 1:
            K, ppairs, pops, pref = parity_sample(args.params, eng, gates, max(1, cores // world), 900 + rank)
            ck_real = T.CloudKey(P, K.offset, K.tv_a, K.tv_b, K.ksk, K.bsk) if rank == 0 else None
            with torch.cuda.stream(stream):
                first_bcast_ms = broadcast_cloud_key(eng, ck_real)      # includes NCCL's lazy communicator set-up
            got = eng.batch_gate_mixed(pops, ppairs)
            parity = {"mismatch_words": int((got != pref).sum()), "gates": gates}

    # ---- bench key: rank 0 uploads + re-lays out, the rest receive one (now warm) NCCL broadcast
    t0 = time.perf_counter()
    ck = synthetic_cloud_key(T, P, 1234) if rank == 0 else None
    key_bcast_ms = None
    if world > 1:
        with torch.cuda.stream(stream):
            key_bcast_ms = broadcast_cloud_key(eng, ck)
    else:
        eng.load_cloud_key(ck)
    key_load_s = time.perf_counter() - t0
    blob_bytes = eng.cloud_key_blob()[1]

    # ---- synthetic ciphertexts (uniform random u32): pinned + pageable host copies, resident copy in HBM
    g = torch.Generator().manual_seed(100 + rank)
    h_in = torch.randint(-2**31, 2**31 - 1, (count, 2, w), dtype=torch.int32, generator=g).pin_memory()
    h_out = torch.empty((count, w), dtype=torch.int32).pin_memory()
    d_in = h_in.to(dev)
    d_out = torch.empty((count, w), dtype=torch.int32, device=dev)
    np_in = h_in.numpy().view(np.uint32)
    np_out = h_out.numpy().view(np.uint32)
    pg_in = np.array(np_in, copy=True)                 # ordinary (pageable) memory, as a caller's Vec is
    pg_out = np.empty((count, w), dtype=np.uint32)

    def step_dev():
        eng.batch_gate_dev("NAND", d_in.data_ptr(), d_out.data_ptr(), count)

    lib = T._load()

    def host_call(src, dst, n):
        rc = lib.tfhe_batch_gate(eng._h, 0, src.ctypes.data_as(C.c_void_p), dst.ctypes.data_as(C.c_void_p), n)
        if rc != 0:
            raise T.EngineError(lib.tfhe_last_error().decode())

    # ---- device-resident leg
    for _ in range(args.warmup):
        step_dev()
    eng.synchronize()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    launches0 = eng.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    br_ms, ks_ms = [], []
    ev0.record(stream)
    for _ in range(args.steps):
        step_dev()
        eng.synchronize()              # also latches the per-kernel event times of this step
        b, k = eng.last_kernel_ms()
        br_ms.append(b); ks_ms.append(k)
    ev1.record(stream)
    barrier()
    launches = eng.kernel_launches - launches0
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None

    # ---- end-to-end legs: host buffers through the C ABI (H2D + kernels + D2H per step)
    def e2e_leg(src, dst):
        for _ in range(min(args.warmup, 2)):
            host_call(src, dst, count)
        barrier()
        t = time.perf_counter()
        for _ in range(args.steps):
            host_call(src, dst, count)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t
        barrier()
        return dt

    e2e_s = e2e_leg(pg_in, pg_out)                      # headline: pageable caller buffers
    e2e_pin_s = e2e_leg(np_in, np_out)
    result_checksum = int(pg_out[:, -1].astype(np.uint64).sum() & 0xFFFFFFFF)
    same_as_pinned = bool(np.array_equal(pg_out, np_out))

    # ---- C5: 1 048 576 mixed gates in total, sharded contiguously over the ranks (strong scaling);
    # device-resident shard processed in blocks of the resident input buffer, CUDA events, max over ranks
    c5_total = 1 << 20
    lo, hi = shard_range(c5_total, rank, world)
    shard = hi - lo
    gops = torch.Generator().manual_seed(0x5EED0005)
    ops_all = torch.randint(0, 6, (c5_total,), dtype=torch.uint8, generator=gops)   # same stream on every rank
    d_ops = ops_all[lo:hi].to(dev)
    c5_ev0, c5_ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    c5_ev0.record(stream)
    done = 0
    c5_sum = 0
    while done < shard:
        nblk = min(count, shard - done)
        lib.tfhe_batch_gate_dev(eng._h, 0, C.c_void_p(d_ops.data_ptr() + done), C.c_void_p(d_in.data_ptr()),
                                C.c_void_p(d_out.data_ptr()), nblk)
        done += nblk
    c5_ev1.record(stream)
    eng.synchronize()
    barrier()
    c5_ms = c5_ev0.elapsed_time(c5_ev1)
    c5_sum = int(d_out[:, -1].to(torch.int64).sum().item() & 0xFFFFFFFF)    # checksum of the shard's last block

    if world > 1:
        tt = torch.tensor([ms_total, e2e_s * 1e3, e2e_pin_s * 1e3, c5_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total, e2e_ms, e2e_pin_ms, c5_ms = (float(x) for x in tt)
        gather = torch.zeros((world, 4), dtype=torch.int64, device=dev)
        mine = torch.tensor([result_checksum, c5_sum, parity["mismatch_words"] if parity["mismatch_words"] is not None else -1,
                             parity["gates"]], dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(gather, mine)
        per_rank = gather.cpu().tolist()
    else:
        e2e_ms, e2e_pin_ms = e2e_s * 1e3, e2e_pin_s * 1e3
        per_rank = [[result_checksum, c5_sum, -1, 0]]

    if rank == 0:
        total = count * world * args.steps
        value = total / (ms_total * 1e-3)
        br_avg = sum(br_ms) / len(br_ms)
        ks_avg = sum(ks_ms) / len(ks_ms)
        flops = flop_per_pbs(P.n, P.l) * count
        fp64_peak = eng.probe_fp64_tflops()
        fp64_peak_3op = eng.probe_fp64_3op_tflops()
        nominal_fp64 = 148 * 64 * 2 * 1.965e9 / 1e12
        achieved = flops / (br_avg * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("blind_rotate_dram_bytes_per_launch")
            except Exception:
                traffic = None
        hbm_peak = 6552.0
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
        # compulsory bytes of one blind-rotation launch: 2 LWE in + extracted sample out per gate
        # plus the Fourier BSK once per launch (it is re-served from L2 after that)
        bsk_bytes = P.n * 2 * P.l * 2 * N * 8
        br_bytes = count * (2 * w * 4 + (N + 1) * 4) + bsk_bytes
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": workload_name(args.params, count), "params": args.params,
                "count_per_gpu": count, "n": P.n, "l": P.l, "bgbit": P.bgbit,
                "l2": f"inputs {count * 2 * w * 4 / 1e6:.0f} MB per step > 126 MB L2 (no flush needed)",
                "inputs": "uniform random u32 LWE pairs and random key material (value-independent data path)",
                "us_per_pbs": ms_total / args.steps * 1e3 / count,
            },
            "e2e": {"value": total / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": int(count * 2 * w * 4),
                    "d2h_bytes_per_step": int(count * w * 4),
                    "api": "tfhe_batch_gate (C ABI) on PAGEABLE host buffers (numpy arrays; what a caller's "
                           "Vec<Ciphertext> is): the engine stages them through its pinned ring",
                    "frac_of_value": total / (e2e_ms * 1e-3) / value,
                    "result_checksum": result_checksum, "same_words_as_pinned_run": same_as_pinned},
            "e2e_pinned": {"value": total / (e2e_pin_ms * 1e-3), "unit": UNIT,
                           "api": "tfhe_batch_gate (C ABI) on pinned host buffers"},
            "gpu_launches": int(launches),
            "kernels_ms_per_step": {"blind_rotate": br_avg, "key_switch": ks_avg},
            "roofline": {
                "kernel": "blind_rotate_kernel_x (+ blind_rotate_kernel_s on a partial last round)", "bound": "fp64",
                "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "frac_of_nominal": achieved / nominal_fp64, "peak_nominal": nominal_fp64,
                "traffic": traffic,
                "peak_source": "DFMA probe kernel measured in this run (tfhe_probe_fp64_tflops); "
                               "MEASURED_PEAKS.json holds no FP64 figure, so the fraction of the nominal "
                               "148 SM x 64 FMA x 2 x 1.965 GHz = 37.2 TFLOP/s is given beside it",
                "algorithmic_flop_per_launch": flops,
                "operand_path": {
                    "peak_three_register_operands": fp64_peak_3op,
                    "frac_of_that": achieved / fp64_peak_3op,
                    "note": "B200 hands its FP64 unit one 64-bit operand per lane per cycle: a DFMA with three distinct "
                            "register operands holds the pipe 3 cycles (measured in this run, "
                            "tfhe_probe_fp64_3op_tflops), with a reused / constant operand 2.2 / 2.06, DADD and "
                            "DMUL 2.  The blind rotation's twiddles and key values differ per lane, so about a "
                            "quarter of its FP64 instructions are of the first kind; its operand-limited ceiling "
                            "is ~0.75 of `peak` (profiles/r2_fp64_operand_probe.json).  `frac` stays against "
                            "`peak`."},
                "hbm_view": {"bound": "hbm", "achieved": br_bytes / (br_avg * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s",
                             "frac": br_bytes / (br_avg * 1e-3) / 1e9 / hbm_peak,
                             "peak_source": hbm_src, "algorithmic_bytes_per_launch": br_bytes},
            },
            "clocks": clocks,
            "key_load_s": key_load_s,
            "key_broadcast": None if world == 1 else {
                "bytes": int(blob_bytes), "warm_ms": key_bcast_ms, "warm_gb_per_s": blob_bytes / (key_bcast_ms * 1e-3) / 1e9,
                "first_ms_incl_communicator_setup": first_bcast_ms,
                "how": "one NCCL broadcast of the re-laid-out key blob (torch.distributed, CUDA events); the first "
                       "broadcast of the process carried the real parity key and pays NCCL's lazy set-up"},
            "c5": {"workload": "1 048 576 gates, op uniform over the 6 batchable gates per element (seed 0x5EED0005), "
                               "contiguous shards over the ranks, device-resident, no per-gate communication",
                   "total_gates": c5_total, "scaling": "strong", "ms": c5_ms, "gates_per_s": c5_total / (c5_ms * 1e-3),
                   "per_rank_last_block_checksum": [r_[1] for r_ in per_rank]},
            "parity_per_rank": [{"rank": i, "real_key_gates": r_[3], "mismatch_words": (None if r_[2] < 0 else r_[2]),
                                 "e2e_result_checksum": r_[0]} for i, r_ in enumerate(per_rank)],
        }
        # secondary kernel: the key switch, an exact u8 one-hot x key-bytes GEMM on tcgen05.mma (< 1 % of the
        # step).  ALGORITHMIC work: the selected rows, N*t*(1 - 2^-basebit) x (n+1) words per gate.
        ks_rows_bytes = N * P.iks_t * (1.0 - 2.0 ** -P.basebit) * w * 4 * count
        line["roofline_key_switch"] = {
            "kernel": "ks_umma_kernel", "bound": "tensor",
            "algorithmic_row_bytes_per_launch": ks_rows_bytes,
            "row_traffic_equivalent_tb_per_s": ks_rows_bytes / (ks_avg * 1e-3) / 1e12,
            "tensor_pipe_active_pct_ncu": 55.4,
            "note": "share of the step < 1 %; the fraction of the tensor roofline is the ncu figure "
                    "sm__pipe_tensor_cycles_active (profiles/r1_ks_umma_ncu_full.json), not an op count: the GEMM "
                    "multiplies by structurally zero one-hot columns, so counted ops are not algorithmic work"}
        if world == 1:
            try:
                line["configs"] = other_configs(T, torch, dev, local, eng)
            except Exception as ex:   # never lose the headline line to a side measurement
                line["configs"] = {"error": repr(ex)}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_leg(args.params, args.cpu_sample, engine=eng)
            line["parity_per_rank"][0]["real_key_gates"] = line["cpu_baseline"].get("parity_checked_gates", 0)
            line["parity_per_rank"][0]["mismatch_words"] = line["cpu_baseline"].get("parity_mismatch_words")
        _emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.close()


if __name__ == "__main__":
    main()

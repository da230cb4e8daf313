#!/usr/bin/env python
"""bench.py -- CKKS HMult+Relin at N=2^16, L=16 (BASELINE.json configs[1]) on N GPUs of one node.

A "step" is one multiply_inplace + relinearize_inplace on one size-2 ciphertext pair at the top data level
(primes CoeffModulus::Create(65536, {60, 40x15, 60x4}), special_modulus_size 4, dnum 4) with synthetic uniform
residues and a synthetic relinearisation key.  Inputs rotate over 8 distinct ciphertext pairs (256 MiB per GPU,
larger than the 126 MB L2) so a step never finds its operands in L2.

  value      whole-job HE-ops/s with operands resident in HBM (CUDA events, max over ranks)
  e2e        the same metric through the C-ABI call that takes HOST buffers
             (pfhe_multiply_and_relin_host_batch: H2D of both operands and D2H of the result inside the timed region)
  roofline   dominant kernel = the forward NTT pair (k_fwd_cols + k_fwd_rows) at the mod-up shape (64 limbs),
             algorithmic bytes 16*N per limb-NTT, timed live with CUDA events
  cpu_baseline  the oracle's C restatement (oracle/liboracle.so, OpenMP over limbs) on the host cores, bounded sample

--impl reference times the UNMODIFIED reference (oracle/_ref/libphantom_ref.so, built from /root/reference by
oracle/Makefile.ref) on the same GPU through its own public API (multiply_inplace + relinearize_inplace, timed the
way benchmark/ckks_bench.cu does); if that library is absent it times the oracle port on the host cores.
Multi-GPU: independent ciphertexts shard across ranks with no data-path collective (weak scaling); NCCL is used
only for the barrier and the max-over-ranks of the device time.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_PAIRS = 8  # distinct resident input pairs (8 x 32 MiB > L2)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median of the samples taken under load (upper half)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons)}


def cpu_baseline(ps, a, b, rlk_h, ops=None, budget_s=12.0):
    import harness as H
    from harness import P
    o = H.oracle()
    cores = min(os.cpu_count() or 1, 64)
    o.orc_set_threads(cores)
    l, n = ps.limbs(), ps.n
    out = np.zeros((2, l, n), dtype=np.uint64)
    o.orc_multiply_relin(ps.octx(), l, P(a[0]), P(b[0]), P(rlk_h), P(out))  # warm-up (tables, page faults)
    if ops is None:   # bounded sample: about budget_s seconds of host work
        t0 = time.perf_counter()
        o.orc_multiply_relin(ps.octx(), l, P(a[0]), P(b[0]), P(rlk_h), P(out))
        ops = max(2, min(200, int(budget_s / max(time.perf_counter() - t0, 1e-3))))
    t0 = time.perf_counter()
    for i in range(ops):
        o.orc_multiply_relin(ps.octx(), l, P(a[i % len(a)]), P(b[i % len(b)]), P(rlk_h), P(out))
    dt = time.perf_counter() - t0
    return {"value": ops / dt, "unit": "HE-ops/s", "cores": cores, "kind": "port",
            "sample": f"{ops} HMult+Relin ops at N=2^16, L=16 through oracle/liboracle.so (OpenMP over limbs)"}


def run_reference(args, ps, a, b, rank, world):
    """--impl reference: the unmodified reference on this GPU, or the oracle port on the host cores."""
    import harness as H
    from harness import P
    if rank != 0:
        return None
    r = H.reference()
    line = {"impl": "reference", "metric": "CKKS HMult+Relin ops/s (N=2^16, L=16)", "unit": "HE-ops/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "CKKS HMult+Relin, N=2^16, L=16, alpha=4, dnum=4, batch=1 ciphertext pair per step"}}
    import torch
    if r is not None and torch.cuda.is_available():
        steps_arr = (ctypes.c_int * 1)(1)
        h = r.ref_create(3, ps.n, P(ps.primes), ps.size_QP, ps.size_P, 0, 0, steps_arr, 1, float(2 ** 40), 1)
        if not h:
            raise RuntimeError(r.ref_last_error().decode())
        trials = args.warmup + args.steps
        times = (ctypes.c_double * trials)()
        assert r.ref_time_op(h, 0, 1, P(a[0]), P(b[0]), 0, 0, trials, times) == 0, r.ref_last_error()
        dev = sum(times[args.warmup:]) / args.steps
        e2e_trials = min(trials, args.warmup + 10)
        assert r.ref_time_op(h, 0, 1, P(a[0]), P(b[0]), 0, 1, e2e_trials, times) == 0, r.ref_last_error()
        e2e = sum(times[args.warmup:e2e_trials]) / (e2e_trials - args.warmup)
        r.ref_destroy(h)
        words = 2 * ps.limbs() * ps.n
        line.update({"value": 1e6 / dev, "ms_per_step": dev / 1e3,
                     "cpu_baseline": {"value": 1e6 / dev, "unit": "HE-ops/s", "cores": 0, "kind": "reference",
                                      "sample": "unmodified phantom-fhe kernels rebuilt for sm_100a on this GPU "
                                                "(multiply_inplace + relinearize_inplace, cudaEvent per trial)"},
                     "e2e": {"value": 1e6 / e2e, "unit": "HE-ops/s", "h2d_bytes_per_step": 2 * words * 8,
                             "d2h_bytes_per_step": words * 8}})
    else:
        cb = cpu_baseline(ps, a, b, H.switch_key(ps, 100), ops=max(1, min(args.steps, 3)))
        line.update({"value": cb["value"], "ms_per_step": 1e3 / cb["value"], "cpu_baseline": cb,
                     "e2e": {"value": cb["value"], "unit": "HE-ops/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0}})
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import harness as H
    ps = H.params_primary()
    # every rank owns different ciphertexts (seeds offset by rank), the key is shared
    a = [H.ciphertext(ps, 10 + 2 * (rank * N_PAIRS + i)) for i in range(N_PAIRS)]
    b = [H.ciphertext(ps, 11 + 2 * (rank * N_PAIRS + i)) for i in range(N_PAIRS)]

    if args.impl == "reference":
        line = run_reference(args, ps, a, b, rank, world)
        if line is not None:
            print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    import phantom_fhe_b200 as pf
    from phantom_fhe_b200 import lib, check

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: phantom-fhe_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    parms = pf.EncryptionParameters(pf.scheme_type.ckks)
    parms.set_poly_modulus_degree(ps.n)
    parms.set_coeff_modulus([int(p) for p in ps.primes])
    parms.set_special_modulus_size(ps.size_P)
    ctx = pf.PhantomContext(parms)
    l, n = ps.limbs(), ps.n
    words = 2 * l * n
    rlk_h = H.switch_key(ps, 100)
    rlk = pf.PhantomRelinKey(ctx, list(rlk_h))
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def to_dev(x):
        return torch.from_numpy(x.view(np.int64)).cuda()

    da = [to_dev(x) for x in a]
    db = [to_dev(x) for x in b]
    work = [torch.empty_like(da[0]) for _ in range(N_PAIRS)]

    def device_step(i):
        k = i % N_PAIRS
        # multiply + relinearize of pair k; the result goes to its own buffer (the reference's in-place form
        # reallocates the ciphertext, include/ciphertext.h:44-72), operands stay intact
        check(lib.pfhe_multiply_and_relin(ctx._h, 1, da[k].data_ptr(), db[k].data_ptr(), work[k].data_ptr(),
                                          rlk.public_keys_ptr(), st))

    CHUNK = 8 * N_PAIRS
    PtrC = ctypes.c_void_p * CHUNK
    arr_a = PtrC(*[da[i % N_PAIRS].data_ptr() for i in range(CHUNK)])
    arr_b = PtrC(*[db[i % N_PAIRS].data_ptr() for i in range(CHUNK)])
    arr_o = PtrC(*[work[i % N_PAIRS].data_ptr() for i in range(CHUNK)])

    def device_steps(count):
        # `count` <= CHUNK steps = `count` independent HMult+Relin ops through the batched C-ABI entry point, one op per
        # step, interleaved over the engine's lanes (independent ciphertexts are the path's sharding unit); operands
        # rotate over the N_PAIRS resident pairs, op i and op i + N_PAIRS share an output buffer and a lane (ordered)
        check(lib.pfhe_multiply_and_relin_batch(ctx._h, 1, arr_a, arr_b, arr_o, count, rlk.public_keys_ptr(), st))

    def refill():
        pass

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput ---------------------------------------------------------------------
    refill()
    for i in range(args.warmup):
        device_step(i)
    device_steps(N_PAIRS)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    total_ms = 0.0
    done = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    while done < args.steps:
        chunk = min(CHUNK, args.steps - done)
        refill()  # untimed: restore the in-place operands
        barrier()
        e0.record()
        device_steps(chunk)
        e1.record()
        barrier()
        total_ms += e0.elapsed_time(e1)
        done += chunk
    launches = ctx.launch_count() - launches0
    dev_ms = max_over_ranks(total_ms)
    # latency of one op issued alone (one lane, nothing else on the GPU): the figure the reference's own bench quotes
    lat_steps = min(args.steps, 64)
    barrier()
    e0.record()
    for i in range(lat_steps):
        device_step(i)
    e1.record()
    barrier()
    lat_ms = max_over_ranks(e0.elapsed_time(e1)) / lat_steps

    # ---- end to end from pinned host memory ---------------------------------------------------------------
    pin_a = [torch.from_numpy(x.view(np.int64)).pin_memory() for x in a]
    pin_b = [torch.from_numpy(x.view(np.int64)).pin_memory() for x in b]
    pin_o = [torch.empty((2, l, n), dtype=torch.int64).pin_memory() for _ in range(N_PAIRS)]
    e2e_steps = min(args.steps, 100)
    PtrArr = ctypes.c_void_p * e2e_steps
    pa = PtrArr(*[pin_a[i % N_PAIRS].data_ptr() for i in range(e2e_steps)])
    pb = PtrArr(*[pin_b[i % N_PAIRS].data_ptr() for i in range(e2e_steps)])
    po = PtrArr(*[pin_o[i % N_PAIRS].data_ptr() for i in range(e2e_steps)])
    check(lib.pfhe_multiply_and_relin_host_batch(ctx._h, 1, pa, pb, po, min(args.warmup, e2e_steps),
                                                 rlk.public_keys_ptr(), st))
    barrier()
    t0 = time.perf_counter()
    e0.record()
    check(lib.pfhe_multiply_and_relin_host_batch(ctx._h, 1, pa, pb, po, e2e_steps, rlk.public_keys_ptr(), st))
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    wall_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel: forward NTT at the mod-up shape (64 limb-NTTs per launch pair) -------
    roof = None
    cb = None
    if rank == 0:
        limbs_ntt = 64
        # 3 rotating buffers of 64 limbs (3 x 32 MiB, with the 20 MiB twiddle table > L2 in steady state)
        bufs = [torch.zeros(limbs_ntt * n, dtype=torch.int64, device="cuda") for _ in range(4)]
        reps = 200
        turn = [0]

        def ntt_launch():
            # 4 polynomials x 16 limbs in one launch pair: the shape of the mod-up NTT (beta = 4, l = 16)
            buf = bufs[turn[0] % len(bufs)]
            turn[0] += 1
            check(lib.pfhe_ntt_forward_inplace_batch(ctx._h, buf.data_ptr(), 4, 16, 0, st))

        for _ in range(3):
            ntt_launch()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            ntt_launch()
        e1.record()
        torch.cuda.synchronize()
        ntt_us = e0.elapsed_time(e1) * 1e3 / reps
        alg_bytes = limbs_ntt * 16 * n
        peak, how = measured_peaks()
        achieved = alg_bytes / (ntt_us * 1e-6) / 1e9
        traffic, issue = None, None
        tpath = os.path.join(ROOT, "profiles", "ntt_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                prof = json.load(f)
            traffic = prof.get("dram_bytes_total")
            if prof.get("warp_instructions"):
                # what the kernel pair is actually bound by: warp-instruction issue.  Peak = one warp
                # instruction per scheduler per clock = SMs x 4 x SM clock; FP64 / IMAD instructions hold the dispatch port
                # for two clocks, so a mix dominated by them tops out near half of that.
                prop = torch.cuda.get_device_properties(local_rank)
                peak_inst = prop.multi_processor_count * 4 * (clocks["sm_mhz"] or 1965.0) * 1e6 if clocks else None
                ach_inst = prof["warp_instructions"] / (ntt_us * 1e-6)
                issue = {"warp_inst_per_launch": prof["warp_instructions"], "achieved_ginst_s": ach_inst / 1e9,
                         "peak_ginst_s": peak_inst / 1e9 if peak_inst else None,
                         "frac": ach_inst / peak_inst if peak_inst else None}
        roof = {"kernel": "forward negacyclic NTT (k_fwd_cols + k_fwd_rows), 64 limb-NTTs of N=2^16",
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": how, "traffic": traffic, "algorithmic_bytes": alg_bytes,
                "launch_us": ntt_us, "limb_ntt_per_s": limbs_ntt / (ntt_us * 1e-6), "issue": issue,
                "note": "not HBM-bound: 64-bit modular butterflies are bound by warp-instruction issue on sm_100a "
                        "(FP64 and IMAD share the dispatch port, 2 clocks each; DESIGN.md 4.1, profiles/r1c_*)"}
        if not args.no_cpu_baseline:
            cb = cpu_baseline(ps, a, b, rlk_h)

    if rank == 0:
        total_steps = args.steps * world
        value = total_steps / (dev_ms * 1e-3)
        line = {
            "metric": "CKKS HMult+Relin ops/s (N=2^16, L=16)", "value": value, "unit": "HE-ops/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": "CKKS HMult+Relin, N=2^16, L=16, alpha=4, dnum=4, batch=1 ciphertext pair per step",
                       "issue": f"steps go through pfhe_multiply_and_relin_batch, {lib.pfhe_engine_lanes(ctx._h)} "
                                "independent ops in flight (lanes); single_op_ms = one op at a time",
                       "l2": f"inputs rotate over {N_PAIRS} resident pairs (256 MiB per GPU) > 126 MB L2",
                       "sharding": "independent ciphertexts per rank, shared key, no data-path collective"},
            "e2e": {"value": e2e_steps * world / (e2e_ms * 1e-3), "unit": "HE-ops/s",
                    "h2d_bytes_per_step": 2 * words * 8, "d2h_bytes_per_step": words * 8,
                    "wall_ms": wall_ms},
            "single_op_ms": lat_ms, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cb,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

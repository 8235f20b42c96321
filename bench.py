#!/usr/bin/env python
"""bench.py — prove_next_layer hot path (batch-STARK proof of one recursion layer) on N B200s.

  python bench.py --gpus N --steps K --warmup W            our CUDA path (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W     the CPU oracle port on the host cores (same metric)

A step = one `prove_all_tables` over one synthetic steady-state recursion layer (tables Const/Public/ALU/Poseidon2/Recompose
at the reference's layer shapes; DEFAULT_FRI = the examples' parameters: log_blowup 2, max_log_arity 2,
log_final_poly_len 5, 54 queries, 15-bit query PoW). Multi-GPU: the path shards over independent proofs (leaves / subtrees
of the 2-to-1 aggregation tree), no data-path collective (weak scaling). `ms_per_layer` is one proof alone on the GPU; `value`
and `e2e` are measured with `--inflight` (default 4) proofs in flight per GPU, one context + stream + host thread each.
Prints ONE JSON line on rank 0 (fd 1 is pointed at stderr for everything else).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

FULL = dict(n_const=1500, n_public=43000, n_alu=60000, n_perms=12000, n_recompose=4000)
# Thread instructions one Poseidon2 permutation costs in k_hash_rows / k_compress (ncu: smsp__inst_executed.sum * 32 /
# permutations of a launch, profiles/r1_ncu_summary.md): the unit conversion of the INT32-pipe roofline below.
INSTR_PER_PERM = {"koala-bear": 5370.0, "baby-bear": 5600.0}   # ncu (koala) / SASS count (baby)
N_SMS, LANES_PER_SM = 148, 128
OUT = sys.stdout
METRIC = "prove_next_layer throughput (layer proofs/s, whole job; ms_per_layer = latency of one proof alone)"


def make_workload(field_name: str, seed: int, scale: float):
    fm = importlib.import_module("plonky3-recursion_b200.field")
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    F = fm.get_field(field_name)
    sizes = {k: max(8, int(v * scale)) for k, v in FULL.items()}
    return F, wl.synthetic_layer(F, seed, min_height=256, **sizes)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device: int):
        super().__init__(daemon=True)
        self.device, self.rows, self._stop_ev = device, [], threading.Event()

    def run(self):
        # NVML in-process (a sample costs ~0.1 ms, so short timed regions still get many samples); nvidia-smi as fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.device)
            bits = [(pynvml.nvmlClocksEventReasonHwSlowdown, 2), (pynvml.nvmlClocksEventReasonHwThermalSlowdown, 3),
                    (pynvml.nvmlClocksEventReasonSwThermalSlowdown, 4), (pynvml.nvmlClocksEventReasonSwPowerCap, 5)]
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop_ev.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                reasons = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                row = [str(sm), str(mx), "", "", "", ""]
                for bit, pos in bits:
                    row[pos] = "Active" if reasons & bit else "Not Active"
                self.rows.append(row)
                self._stop_ev.wait(0.005)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=5)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) >= 6 and r[2 + i].lower() == "active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def run_reference(args):
    """CPU arm: the oracle port (kind "port": the Rust reference cannot be built here) with all host threads, on a bounded
    1/16-scale sample of the same layer; value is extrapolated linearly in table rows to the full layer."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use every host core
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from common import make_oracle
    lib = importlib.import_module("plonky3-recursion_b200.lib")
    sample = 1.0 / 16
    F, L = make_workload(args.field, 1, sample)
    orc = make_oracle(args.field, lib.DEFAULT_FRI)
    for _ in range(args.warmup):
        orc.prove(L.insts, L.preps, L.traces, L.pubs)
    t0 = time.time()
    for _ in range(args.steps):
        orc.prove(L.insts, L.preps, L.traces, L.pubs)
    dt = (time.time() - t0) / args.steps
    value = sample / dt
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 (31-bit Montgomery field, degree-4 extension)", "data": "synthetic",
        "config": {"workload": f"synthetic steady-state recursion layer ({args.field}), 1/16-scale sample per step", "shapes": L.shapes},
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port",
                         "sample": "1/16-scale layer per step (rows/16 per table), oracle/liboracle.so with OpenMP; "
                                   "value = (1/16)/seconds, i.e. extrapolated linearly in rows to the full layer"},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=OUT, flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = importlib.import_module("plonky3-recursion_b200.lib")
    ctx = lib.Context(args.field, lib.DEFAULT_FRI, device=local)
    F, L = make_workload(args.field, 1 + rank, args.scale)
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    prover = lib.BatchStarkProver(ctx, pinned_output=True)
    tb_res = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)       # device-resident inputs -> `value`
    # pinned host inputs -> `e2e`: row-major matrices for Const / Public / Recompose as the reference's trace builders leave
    # them; the Poseidon2 and ALU tables go in as operation lists and are expanded on the device (p3r_prove_ops)
    tb_pin = lib.TraceBatch(ctx, L.traces, L.pubs, pinned=True, p2_ops=L.p2_ops, alu_ops=L.alu_ops)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    def timed_steps(fn, steps):
        """fn() once per step, each step timed with CUDA events on the prover's stream; L2 flushed between steps (untimed)."""
        total = 0.0
        for _ in range(steps):
            flush_l2()
            ctx.timer_start()
            fn()
            total += ctx.timer_stop()
        return total

    # ---- warm-up + per-class breakdown (untimed) ----
    for _ in range(max(args.warmup, 3)):
        prover.prove_resident(tb_res, pd, copy=False)
    ctx.reset_kernel_stats()
    ctx.set_kernel_timing(lib.KERNEL_CLASSES)
    for _ in range(2):
        prover.prove_resident(tb_res, pd, copy=False)
    breakdown = {k: v["ms"] / 2 for k, v in ctx.kernel_stats().items()}
    dominant = max(breakdown, key=breakdown.get)
    # live timing, inside the timed region, of the dominant kernel class and of the LDE (the HBM-roofline kernel)
    ctx.set_kernel_timing(sorted({dominant, "ntt_lde"}))
    ctx.reset_kernel_stats()

    # ---- timed region A: device-resident inputs ----
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = ctx.launch_count()
    t_res = timed_steps(lambda: prover.prove_resident(tb_res, pd, copy=False), args.steps)
    launches = ctx.launch_count() - l0
    barrier()
    kstats = ctx.kernel_stats()
    ks, ks_lde = kstats[dominant], kstats["ntt_lde"]
    ctx.set_kernel_timing([])
    proof_words = prover.last_proof_words
    # ---- proofs in flight: one context (stream, arena) + one host thread per concurrent proof -------------------------
    # A single proof leaves the GPU idle during its latency-bound parts (small Merkle levels, host hand-offs); a prover that
    # serves an aggregation tree always has independent proofs, so `value` is measured with `--inflight` proofs per batch.
    lanes = [(ctx, pd, prover, tb_res, tb_pin)]
    # `ProverData::from_airs_and_degrees` (preprocessed LDE + MMCS tree + programs, once per circuit shape,
    # host matrices in): wall clock of a warm C-ABI call, reported beside the per-layer numbers, not part of `value`.
    pd_again = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    prep_commit_ms = pd_again.commit_ms           # p3r_prep_commit alone: Montgomery host matrices in, cap on the host out
    pd_again.close()
    for _ in range(1, args.inflight):
        c2 = lib.Context(args.field, lib.DEFAULT_FRI, device=local)
        pd2 = lib.ProverData.from_airs_and_degrees(c2, L.insts, L.preps)
        lanes.append((c2, pd2, lib.BatchStarkProver(c2, pinned_output=True), lib.TraceBatch(c2, L.traces, L.pubs).upload(pd2),
                      lib.TraceBatch(c2, L.traces, L.pubs, pinned=True, p2_ops=L.p2_ops, alu_ops=L.alu_ops)))

    def batch_steps(e2e, steps):
        """`steps` batches of len(lanes) concurrent proofs. Per batch: L2 flush (untimed), a start event on every stream, the
        proofs (one host thread each), a stop event per stream; batch time = max over streams. Returns (ms total, launches)."""
        n = len(lanes)
        go, done = threading.Barrier(n + 1), threading.Barrier(n + 1)
        ms = [0.0] * n
        stop = [False]

        def worker(k):
            c, p, pr, tr, tp = lanes[k]
            while True:
                go.wait()
                if stop[0]:
                    return
                if e2e:
                    pr.prove_all_tables(tp, p)
                else:
                    pr.prove_resident(tr, p, copy=False)
                ms[k] = c.timer_stop()
                done.wait()

        ths = [threading.Thread(target=worker, args=(k,), daemon=True) for k in range(n)]
        for t in ths:
            t.start()
        total, l0 = 0.0, sum(ln[0].launch_count() for ln in lanes)
        for _ in range(steps):
            flush_l2()
            for ln in lanes:
                ln[0].timer_start()
            go.wait()
            done.wait()
            total += max(ms)
        stop[0] = True
        go.wait()
        for t in ths:
            t.join()
        return total, sum(ln[0].launch_count() for ln in lanes) - l0

    # Host wait mode of the prover threads in the throughput regions (p3r_set_wait_mode; the library default is yield).
    cores = len(os.sched_getaffinity(0))
    wait_mode = args.wait or os.environ.get("P3R_WAIT", "yield")
    ctx.set_wait_mode(wait_mode)
    batch_steps(False, max(args.warmup, 3))
    barrier()
    t_batch, launches_batch = batch_steps(False, args.steps)
    barrier()
    # ---- end to end through the C ABI with host buffers (H2D of traces + D2H of the proof inside the timed region) ----
    batch_steps(True, 2)
    barrier()
    t_e2e, _ = batch_steps(True, args.steps)
    barrier()
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([t_res, t_e2e, t_batch], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_res, t_e2e, t_batch = float(t[0]), float(t[1]), float(t[2])
    if rank == 0:
        peak, peak_kind = peaks()
        ms_layer = t_res / args.steps                       # latency of ONE proof alone on the GPU
        ms_step = t_batch / args.steps                      # one batch = args.inflight concurrent proofs
        value = world * args.inflight * args.steps / (t_batch / 1e3)
        e2e_value = world * args.inflight * args.steps / (t_e2e / 1e3)
        traffic = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f)
        sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6

        def hbm_roofline(name, st):
            ach = (st["bytes"] / 1e9) / (st["ms"] / 1e3) if st["ms"] > 0 else 0.0
            return {"kernel": name, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic.get(name), "traffic_scope": traffic.get(name + "_scope"), "peak_source": peak_kind,
                    "algorithmic_bytes_per_step": st["bytes"] / args.steps,
                    "kernel_ms_per_step": st["ms"] / args.steps, "launches_per_step": st["launches"] / args.steps}

        if dominant in ("hash_rows", "compress"):
            # Poseidon2 is bound by the integer pipes (31-bit Montgomery arithmetic, no tensor cores): achieved = thread
            # instructions per second of the class, peak = SMs x 128 lanes x SM clock sampled during the run.
            perms_s = ks["perms"] / (ks["ms"] / 1e3) if ks["ms"] > 0 else 0.0
            ach = perms_s * INSTR_PER_PERM[args.field] / 1e12
            pk = N_SMS * LANES_PER_SM * sm_hz / 1e12
            roofline = {"kernel": dominant, "bound": "int32_pipe", "achieved": ach, "peak": pk, "unit": "Tlane-instr/s",
                        "frac": ach / pk, "traffic": traffic.get(dominant), "traffic_scope": traffic.get(dominant + "_scope"),
                        "peak_source": "148 SMs x 128 lanes x sampled SM clock",
                        "permutations_per_s": perms_s, "instr_per_permutation": INSTR_PER_PERM[args.field],
                        "permutations_per_step": ks["perms"] / args.steps, "kernel_ms_per_step": ks["ms"] / args.steps,
                        "launches_per_step": ks["launches"] / args.steps,
                        "algorithmic_bytes_per_step": ks["bytes"] / args.steps,
                        "hbm_frac": (ks["bytes"] / 1e9) / (ks["ms"] / 1e3) / peak if ks["ms"] > 0 else 0.0}
        else:
            roofline = hbm_roofline(dominant, ks)
        line = {
            "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "ms_per_layer": ms_layer, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (31-bit Montgomery field, degree-4 extension)", "data": "synthetic",
            "config": {"workload": f"synthetic steady-state recursion layer ({args.field}, recursive_fibonacci layer shape), "
                                   f"one proof per GPU per step", "shapes": L.shapes, "fri": lib.DEFAULT_FRI, "scale": args.scale,
                       "l2": "flushed between timed steps (256 MiB fill)",
                       "parallelism": f"independent proofs x{world} GPUs, {args.inflight} proofs in flight per GPU and step",
                       "proofs_per_step": world * args.inflight, "host_wait": wait_mode, "host_cores": cores,
                       "single_proof_latency_ms": ms_layer, "single_proof_proofs_per_s": world * args.steps / (t_res / 1e3),
                       "proof_words": proof_words, "prep_commit_ms": prep_commit_ms},
            "e2e": {"value": e2e_value, "unit": "proofs/s", "ms_per_step": t_e2e / args.steps,
                    "h2d_bytes_per_step": tb_pin.h2d_bytes * args.inflight, "d2h_bytes_per_step": proof_words * 4 * args.inflight},
            "gpu_launches": launches_batch,
            "roofline": roofline,
            "roofline_lde": hbm_roofline("ntt_lde", ks_lde),
            "kernel_breakdown_ms": {k: round(v, 4) for k, v in breakdown.items()},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            from common import make_oracle
            sample = 1.0 / 8
            Fc, Lc = make_workload(args.field, 1, sample)
            orc = make_oracle(args.field, lib.DEFAULT_FRI)
            orc.prove(Lc.insts, Lc.preps, Lc.traces, Lc.pubs)
            t0 = time.time()
            reps = 2
            for _ in range(reps):
                orc.prove(Lc.insts, Lc.preps, Lc.traces, Lc.pubs)
            dt = (time.time() - t0) / reps
            line["cpu_baseline"] = {"value": sample / dt, "unit": "proofs/s", "cores": os.cpu_count() or 1, "kind": "port",
                                    "sample": f"1/8-scale layer (rows/8 per table) proved {reps}x by oracle/liboracle.so (OpenMP), "
                                              f"{dt:.2f} s each; value extrapolated linearly in rows to the full layer"}
        print(json.dumps(line), file=OUT, flush=True)
    for c_, p_, _, tr_, _ in lanes:
        tr_.close()
        p_.close()
        c_.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--field", default="koala-bear", choices=["koala-bear", "baby-bear"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wait", choices=["spin", "yield", "block"], default=None,
                    help="host wait mode in the throughput regions (default: the library default, yield)")
    ap.add_argument("--inflight", type=int, default=4, help="concurrent proofs per GPU in the throughput regions")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: keep the real stdout aside and point fd 1 at stderr, so that anything a library
    # prints there (NCCL's version banner, for one) cannot precede it.
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        if args.steps == 20:
            args.steps = 3
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

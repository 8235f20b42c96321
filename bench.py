#!/usr/bin/env python
"""bench.py — the prove_next_layer / aggregation-layer hot path (batch-STARK proofs of recursion layers) on N B200s.

  python bench.py --gpus N --steps K --warmup W             our CUDA path (one rank per GPU under torchrun for N > 1)
  python bench.py --impl reference --steps K --warmup W      the CPU oracle port on the host cores (same metric)

BASELINE.json's metric has two halves and one command line serves both, so the line carries both:
  * `value` / `e2e` = AGGREGATION PROOFS PER SECOND of the 2-to-1 aggregation tree (recursion/examples/
    recursive_aggregation.rs:624-707: 8 base proofs -> 4 -> 2 -> 1) at every N — the same quantity at N = 1, 2, 4, 8, so the
    driver's scaling arithmetic compares like with like. Many 8-leaf trees are in flight; on N > 1 ranks the nodes of a tree
    live on different GPUs (block partition of the leaves, rotated per tree so every rank proves the same number of nodes) and
    every child proof whose parent lives elsewhere is handed over INSIDE the timed region by NCCL send/recv (pinned host ->
    device -> NVLink -> device -> pinned host; the parent's witness generation is host code). A node's Public table carries a
    checksum of each child proof, so a lost or mis-routed proof changes the root; `tree.roots_checksum` is the same number at
    every N. `value`: every proof starts from device-resident traces (only the child-dependent rows are written per proof);
    `e2e`: every proof goes through the reference-facing call with HOST buffers (row-major matrices / operation lists in
    pinned memory copied to the device per proof, the proof copied back).
  * `ms_per_layer` = latency of ONE prove_next_layer-shaped proof alone on one GPU (BASELINE's first half), L2 flushed between
    steps, with the per-kernel-class breakdown and the roofline of the dominant kernel class.
A step = `trees_per_step` trees (2 per GPU) = 7 aggregation proofs + 8 leaf proofs each. Tasks and hand-offs follow
`aggregation.order_key` (default "block": trees in blocks of eight, higher levels first inside a block, so every rank proves
runs of same-shaped proofs at any N). At N = 1 the line also carries `cpu_baseline` (the oracle on the host cores, full-size
node layer) and `parity` (the GPU proof of that layer == the oracle prover's, all words, checked outside the timed regions;
a mismatch fails the run). `tree.per_rank`, `tree.proof_ms_by_shape_rank0` and `tree.node_proofs_in_flight_hist_rank0`
are diagnostics of the multi-GPU pipeline.
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

# Synthetic steady-state recursion layer (estimates from the book, SURVEY.md Appendix B9 is open): operation counts per table.
FULL = dict(n_const=1500, n_public=43000, n_alu=60000, n_perms=12000, n_recompose=4000)
# Thread instructions one Poseidon2 permutation costs in k_hash_rows / k_compress (ncu: smsp__inst_executed.sum * 32 /
# permutations of a launch, profiles/): the unit conversion of the INT32-pipe roofline below.
INSTR_PER_PERM = {"koala-bear": 5370.0, "baby-bear": 5600.0}
N_SMS, LANES_PER_SM = 148, 128
OUT = sys.stdout
METRIC = ("aggregation proofs/s (2-to-1 tree over 8 base proofs, many trees in flight, child proofs handed over inside the timed "
          "region); prove_next_layer ms/layer (KoalaBear) of one proof alone in ms_per_layer")
PUBLIC_INST = 1          # instance order [Const, Public, ALU, Poseidon2, Recompose]
PATCH_WORDS = 16         # two 8-word child-proof checksums (nodes) / (tree, leaf) identity (leaves)


def fri_params(lib, log_final_poly_len):
    d = dict(lib.DEFAULT_FRI)
    d["log_final_poly_len"] = log_final_poly_len
    return d


def make_workload(field_name: str, seed: int, scale: float, min_height: int = 256, **packing):
    fm = importlib.import_module("plonky3-recursion_b200.field")
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    F = fm.get_field(field_name)
    sizes = {k: max(8, int(v * scale)) for k, v in FULL.items()}
    return F, wl.synthetic_layer(F, seed, min_height=min_height, **sizes, **packing)


def tree_workloads(field_name: str, scale: float):
    """The three circuit shapes of the aggregation example (recursive_aggregation.rs:632,666; log_final_poly_len 6 => minimum
    table height 2^(6+2+1) = 512, batch_stark_prover/packing.rs:100-106):
      leaf  — a base proof: the extension-degree-1 base circuit (TablePacking::new(1, 1), packed Horner depth 2);
      l1    — aggregation level 1: verifier circuit of two small base proofs, TablePacking::new(2, 2), Horner depth 2
              (synthetic layer at half the steady-state size);
      node  — aggregation levels >= 2: verifier circuit of two recursion proofs = the steady-state layer, packing (1, 3, k=4)."""
    fm = importlib.import_module("plonky3-recursion_b200.field")
    wl = importlib.import_module("plonky3-recursion_b200.workload")
    F = fm.get_field(field_name)
    _, node = make_workload(field_name, 1, scale, 512)
    _, l1 = make_workload(field_name, 2, 0.5 * scale, 512, public_lanes=2, alu_lanes=2, horner_k=2)
    leaf = wl.base_layer_fibonacci(F, 1000, min_height=512)
    return F, {"leaf": leaf, "l1": l1, "node": node}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


def proof_checksum(proof: np.ndarray, p: int) -> np.ndarray:
    """8 words: word i = sum of proof[i::8] mod p. What a parent's Public table carries of each child proof."""
    n = proof.size - proof.size % 8
    s = proof[:n].reshape(-1, 8).astype(np.uint64).sum(axis=0)
    if n < proof.size:
        s[: proof.size - n] += proof[n:].astype(np.uint64)
    return (s % np.uint64(p)).astype(np.uint32)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, device: int, period_s: float = 0.005):
        super().__init__(daemon=True)
        self.device, self.rows, self._stop_ev, self.period_s = device, [], threading.Event(), period_s

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
                self._stop_ev.wait(self.period_s)
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


# ----------------------------------------------------------------------------------------------------------------------------
# CPU arm
# ----------------------------------------------------------------------------------------------------------------------------
def time_oracle(field, fri, L, steps, warmup, budget_s):
    """Times oracle/liboracle.so (OpenMP, all host cores) on the FULL-SIZE layer `L`: `steps` proofs after `warmup`, inputs
    marshalled once outside the loop. If the requested count cannot finish inside `budget_s`, fewer are run and the line says
    so (never an extrapolation in size). Returns (seconds per proof, steps run, warm-ups run)."""
    from common import make_oracle
    orc = make_oracle(field, fri)
    orc.set_threads(len(os.sched_getaffinity(0)))   # OMP_NUM_THREADS may have been read (as 1, under torchrun) long ago
    orc.set_fast_paths(True)
    run = orc.prepare(L.insts, L.preps, L.traces, L.pubs)
    t0 = time.perf_counter()
    run()
    first = time.perf_counter() - t0
    w_done = 1
    while w_done < warmup and (w_done + 1 + steps) * first < budget_s:
        run()
        w_done += 1
    n = max(1, min(steps, int(budget_s / first) - w_done))
    t0 = time.perf_counter()
    last = None
    for _ in range(n):
        last = run()
    dt = (time.perf_counter() - t0) / n
    time_oracle.last_proof = np.array(last, copy=True)    # bench.py compares the GPU proof of the same inputs with it
    return dt, n, w_done


def run_reference(args):
    """CPU arm: the oracle port (kind "port": the Rust reference cannot be built in this image — no cargo / rustc) with all
    host threads, on the full-size aggregation-node layer, one proof per step. The published Rust number (other hardware) is
    quoted beside it. The oracle runs its CPU-arm routes here (oracle/fast_paths.inc: AVX2 / AVX-512 Poseidon2, eight-column
    strips interpolated once and shared by the LDE, the quotient domain and the openings, an eight-row constraint interpreter,
    batched inversions, OpenMP over all host cores) — the layout of a packed-field CPU prover, still a port and not the tuned
    Rust code, so the driver's ratio to this line remains an upper bound on the speed-up over the real reference."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for its workers; the CPU arm is meant to use every host core
    cores = len(os.sched_getaffinity(0))
    os.environ["OMP_NUM_THREADS"] = str(cores)
    lib = importlib.import_module("plonky3-recursion_b200.lib")
    fri = fri_params(lib, 6)
    F, L = make_workload(args.field, 1, args.scale, 512)
    dt, n, w = time_oracle(args.field, fri, L, args.steps, args.warmup, args.cpu_budget_s)
    value = 1.0 / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "proofs/s", "n_gpus": args.gpus, "steps": n,
        "warmup": w, "ms_per_step": dt * 1e3, "ms_per_layer": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 (31-bit field, degree-4 extension; Montgomery AVX2/AVX-512 lanes on the CPU)", "data": "synthetic",
        "steps_requested": args.steps, "same_steps": n == args.steps, "same_config": args.scale == 1.0, "extrapolated": False,
        "config": {"workload": f"one full-size aggregation-node layer proof per step ({args.field}, scale {args.scale}): the CPU arm "
                               "is charged only the level >= 2 node proofs; the leaf and level-1 proofs the GPU arm also proves "
                               "inside its timed region are free here (favours the CPU)",
                   "shapes": L.shapes, "fri": fri, "published_reference": "109 ms/layer, Apple M4 Pro 14 cores "
                   "(book/src/appendix/benchmark.md:55,60) — other hardware, the Rust prover itself"},
        "cpu_baseline": {"value": value, "unit": "proofs/s", "cores": cores, "kind": "port",
                         "sample": f"{n} full-size layer proofs ({dt:.2f} s each) by oracle/liboracle.so with OpenMP on {cores} "
                                   "host threads; marshalling outside the timed loop; no size extrapolation"},
        "e2e": {"value": value, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=OUT, flush=True)


# ----------------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------------
class Lane:
    """One proving context (stream, arena, host thread) with the three circuit shapes of the tree prepared on it."""

    def __init__(self, lib, field, fri, device, shapes, priority_shapes=()):
        # the shapes named in `priority_shapes` are proved on a second context of this lane whose streams have high CUDA
        # priority (p3r_ctx_set_stream_priority): the short dependent kernels of a small proof do not queue behind the
        # pending blocks of the full-size proofs of the other lanes
        self.ctxs = [lib.Context(field, fri, device=device)]
        provers = [lib.BatchStarkProver(self.ctxs[0], pinned_output=True)]
        if any(n in priority_shapes for n in shapes):
            self.ctxs.append(lib.Context(field, fri, device=device))
            self.ctxs[1].set_stream_priority(True)
            provers.append(lib.BatchStarkProver(self.ctxs[1], pinned_output=True))
        self.kind = {}
        self.log = []                     # (shape, start, end) of every proof of this lane, host clock
        for name, L in shapes.items():
            which = 1 if name in priority_shapes else 0
            ctx = self.ctxs[which]
            pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
            res = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
            host = lib.TraceBatch(ctx, L.traces, L.pubs, pinned=True, p2_ops=L.p2_ops, alu_ops=L.alu_ops)
            w = L.insts[PUBLIC_INST].main_width
            rows = max(1, PATCH_WORDS // w)
            # the patched rows must be padding of the Public table: index 0, multiplicity 0 in its preprocessed columns, so
            # that their values are free (the WitnessChecks bus ignores them) while still being committed
            if np.any(L.preps[PUBLIC_INST][-rows:] != 0) or np.any(L.traces[PUBLIC_INST][-rows:] != 0):
                raise ValueError(f"shape '{name}': the last {rows} rows of the Public table are not padding")
            self.kind[name] = (pd, res, host, (1 << L.insts[PUBLIC_INST].log_height) - rows, rows, w, provers[which])

    def prove(self, name, patch_words, host_buffers: bool):
        """One proof of shape `name` whose Public table carries `patch_words` in its last (padding, multiplicity-0) rows."""
        pd, res, host, row0, rows, w, prover = self.kind[name]
        tb = host if host_buffers else res
        t0 = time.perf_counter()
        tb.write_rows(pd, PUBLIC_INST, row0, np.asarray(patch_words, dtype=np.uint32)[: rows * w].reshape(rows, w))
        if host_buffers:
            out = prover.prove_all_tables(tb, pd, copy=False).copy()
        else:
            out = prover.prove_resident(tb, pd, copy=False).copy()
        self.log.append((name, t0, time.perf_counter()))
        return out

    def close(self):
        for pd, res, host, *_ in self.kind.values():
            res.close()
            pd.close()
        for c in self.ctxs:
            c.close()

    def launch_count(self):
        return sum(c.launch_count() for c in self.ctxs)

    def timer_start(self):
        for c in self.ctxs:
            c.timer_start()

    def timer_stop(self):
        return max(c.timer_stop() for c in self.ctxs)


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if world > 1 and args.pin_cores:
        # one contiguous slice of the host cores per rank: the ranks' polling and launching threads stop migrating onto each other
        avail = sorted(os.sched_getaffinity(0))
        per = len(avail) // world
        if per >= 1:
            os.sched_setaffinity(0, avail[local * per:(local + 1) * per])
    lib = importlib.import_module("plonky3-recursion_b200.lib")
    agg = importlib.import_module("plonky3-recursion_b200.aggregation")
    F, shapes = tree_workloads(args.field, args.scale)
    p = F.p
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def flush_l2():
        flush.fill_(1)
        torch.cuda.synchronize()

    # ======== part 1: latency of ONE prove_next_layer proof alone (recursive_fibonacci parameters: log_final_poly_len 5) ========
    ctx = lib.Context(args.field, lib.DEFAULT_FRI, device=local)
    # host wait of the prover threads: pure yield-polling while every proving thread can have a core of its own, the sleeping
    # poll once the node's proving threads (ranks x lanes) reach the number of host cores (p3r_set_wait_mode)
    cores_avail = len(os.sched_getaffinity(0)) * (world if (world > 1 and args.pin_cores) else 1)
    wait_mode = args.wait or os.environ.get("P3R_WAIT") or ("sleep" if world * args.inflight >= cores_avail else "yield")
    ctx.set_wait_mode(wait_mode)
    L = shapes["node"]
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    prep_commit_ms = pd.commit_ms
    prover = lib.BatchStarkProver(ctx, pinned_output=True)
    tb_res = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
    for _ in range(max(args.warmup, 3)):
        prover.prove_resident(tb_res, pd, copy=False)
    ctx.reset_kernel_stats()
    ctx.set_kernel_timing(lib.KERNEL_CLASSES)
    for _ in range(2):
        prover.prove_resident(tb_res, pd, copy=False)
    breakdown = {k: v["ms"] / 2 for k, v in ctx.kernel_stats().items()}
    launches_per_proof = sum(v["launches"] for v in ctx.kernel_stats().values()) / 2
    dominant = max(breakdown, key=breakdown.get)
    # live timing, inside the timed region, of the dominant kernel class and of the LDE (the HBM-roofline kernel)
    ctx.set_kernel_timing(sorted({dominant, "ntt_lde"}))
    ctx.reset_kernel_stats()
    sampler = ClockSampler(local, args.clock_sample_ms / 1e3)
    sampler.start()
    barrier()
    t_res = 0.0
    H_pub = 1 << L.insts[PUBLIC_INST].log_height
    for step in range(args.steps):
        # a different proof every step (the Public table's padding rows carry the step number), so that challenges, openings
        # and every challenge-dependent descriptor differ from the previous proof as they do in a real chain of layers
        tb_res.write_rows(pd, PUBLIC_INST, H_pub - 1, np.full((1, L.insts[PUBLIC_INST].main_width), step + 1, dtype=np.uint32))
        flush_l2()
        ctx.timer_start()
        prover.prove_resident(tb_res, pd, copy=False)
        t_res += ctx.timer_stop()
    barrier()
    kstats = ctx.kernel_stats()
    ks, ks_lde = kstats[dominant], kstats["ntt_lde"]
    ctx.set_kernel_timing([])
    proof_words = prover.last_proof_words
    tb_res.close()
    pd.close()

    # ======== part 2: the aggregation tree ========
    prio = tuple(x for x in args.priority_shapes.split(",") if x)
    lanes = [Lane(lib, args.field, fri_params(lib, 6), local, shapes, prio) for _ in range(args.inflight)]
    kind_of_level = lambda lvl: "leaf" if lvl == 0 else ("l1" if lvl == 1 else "node")
    depth = args.leaves.bit_length() - 1
    sizes = {}
    for lvl in range(depth + 1):       # proof sizes per level (shape-determined, identical on every rank) + warm-up of every shape
        sizes[lvl] = int(lanes[0].prove(kind_of_level(lvl), np.zeros(PATCH_WORDS, dtype=np.uint32), False).size)
    transport = None
    if world > 1:
        transport = agg.TorchTransport(dist, torch, dev, agg.MSG_HEADER_WORDS + max(sizes.values()), max_outstanding=48)
    mode = {"host": False}

    region_base = [0]     # first tree of the current region: a leaf's identity is its tree's index WITHIN the region, so the
                          # k-th timed tree is the same tree (same root proof) at every N

    def prove_leaf(k, t, i):
        ident = np.zeros(PATCH_WORDS, dtype=np.uint32)
        ident[0], ident[1] = (t - region_base[0]) % p, i
        return lanes[k].prove("leaf", ident, mode["host"])

    def prove_node(k, t, nd, left, right):
        patch = np.concatenate([proof_checksum(left, p), proof_checksum(right, p)])
        return lanes[k].prove(kind_of_level(nd.level), patch, mode["host"])

    ex = agg.TreeExecutor(rank, world, args.leaves, args.inflight, prove_leaf, prove_node, transport, sizes, skew=args.skew,
                          timeout_s=args.tree_timeout_s, comm_poll_s=args.comm_poll_us / 1e6, order=args.order)
    trees_per_step = args.trees_per_step or 2 * world
    n_agg = args.leaves - 1
    next_tree = [0]

    def tree_region(n_trees, host_buffers):
        """n_trees trees, no barrier between them; CUDA events on every lane's stream, time = max over lanes."""
        mode["host"] = host_buffers
        region_base[0] = next_tree[0]
        barrier()
        l0 = sum(ln.launch_count() for ln in lanes)
        for ln in lanes:
            ln.log.clear()
            ln.timer_start()
        w0 = time.perf_counter()
        c0 = time.process_time()
        out = ex.run(n_trees, first_tree=next_tree[0])
        ms = max(ln.timer_stop() for ln in lanes)
        out["wall_ms"] = (time.perf_counter() - w0) * 1e3
        out["cpu_s"] = time.process_time() - c0       # user + system time of every thread of this rank
        out["launches"] = sum(ln.launch_count() for ln in lanes) - l0
        next_tree[0] += n_trees
        barrier()
        out["ms"] = ms
        # host-clock duration of the proofs by shape, and how many full-size node proofs were in flight at once (time-weighted)
        ev, by = [], {}
        for ln in lanes:
            for name, a, b in ln.log:
                by.setdefault(name, []).append(b - a)
                if name == "node":
                    ev += [(a, 1), (b, -1)]
        out["proof_ms"] = {k: round(1e3 * float(np.mean(v)), 3) for k, v in by.items()}
        hist, cur, last = [0.0] * (len(lanes) + 1), 0, None
        for tt, d in sorted(ev):
            if last is not None:
                hist[cur] += tt - last
            cur, last = cur + d, tt
        tot = sum(hist) or 1.0
        out["nodes_in_flight"] = [round(h / tot, 3) for h in hist]
        return out

    warm_trees = max(world, 2)           # every rank plays every role once: all NCCL peer channels are set up
    tree_region(warm_trees, False)
    tree_region(warm_trees, True)
    n_trees = args.steps * trees_per_step
    first_timed = next_tree[0]
    r_res = tree_region(n_trees, False)
    next_tree[0] = first_timed           # same trees again through host buffers: same proofs, same roots
    r_e2e = tree_region(n_trees, True)
    clocks = sampler.stop()

    common_trees = 2 * args.steps        # the trees every N proves (N = 1 proves exactly these): comparable across N

    def roots_sum(r):
        s = 0
        for t, pr in r["roots"].items():
            if t - first_timed < common_trees:
                s += int(proof_checksum(pr, p).astype(np.uint64).sum()) * (2 * (t - first_timed) + 1)
        return s
    stats = torch.tensor([t_res, r_res["ms"], r_e2e["ms"], r_res["wall_ms"], r_e2e["wall_ms"],
                          sum(r_res["idle_s"]) / len(lanes) / (r_res["wall_ms"] / 1e3),
                          r_res["cpu_s"] / (r_res["wall_ms"] / 1e3)], dtype=torch.float64, device=dev)
    sums = torch.tensor([r_res["sent_bytes"], r_e2e["sent_bytes"], r_res["launches"], r_e2e["launches"],
                         roots_sum(r_res) % (1 << 59), roots_sum(r_e2e) % (1 << 59),
                         sum(r_res["proved"].values()), int(1e6 * sum(r_res["idle_s"]) / len(lanes))], dtype=torch.int64, device=dev)
    mine = torch.tensor([r_res["ms"], r_e2e["ms"], float(clocks.get("sm_mhz") or 0.0), r_res["cpu_s"] / (r_res["wall_ms"] / 1e3),
                         sum(r_res["idle_s"]) / len(lanes) / (r_res["wall_ms"] / 1e3), float(sum(r_res["proved"].values()))],
                        dtype=torch.float64, device=dev)
    per_rank = [mine]
    if world > 1:
        per_rank = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(per_rank, mine)
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    per_rank = [[round(float(x), 3) for x in t] for t in per_rank]
    t_res, ms_res, ms_e2e, wall_res, wall_e2e, idle_worst, host_cores_busy = (float(x) for x in stats)
    sent_res, sent_e2e, launches_res, launches_e2e, rsum_res, rsum_e2e, proved_total, idle_us = (int(x) for x in sums)

    if rank == 0:
        peak, peak_kind = peaks()
        ms_layer = t_res / args.steps
        value = n_agg * n_trees / (ms_res / 1e3)
        e2e_value = n_agg * n_trees / (ms_e2e / 1e3)
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
                        "peak_source": "148 SMs x 128 lanes x sampled SM clock", "scope": "one proof alone (part 1 of the run)",
                        "permutations_per_s": perms_s, "instr_per_permutation": INSTR_PER_PERM[args.field],
                        "permutations_per_step": ks["perms"] / args.steps, "kernel_ms_per_step": ks["ms"] / args.steps,
                        "launches_per_step": ks["launches"] / args.steps,
                        "algorithmic_bytes_per_step": ks["bytes"] / args.steps,
                        "hbm_frac": (ks["bytes"] / 1e9) / (ks["ms"] / 1e3) / peak if ks["ms"] > 0 else 0.0}
        else:
            roofline = hbm_roofline(dominant, ks)
        lde = hbm_roofline("ntt_lde", ks_lde)
        # second bound of the LDE: the integer-multiplier pipe. 12.1 Montgomery products / clk / SM measured
        # (profiles/r1_ubench.txt); per input element (1+B)*log2(n)/2 butterflies at ~4/3 products each (derived twiddles) and
        # 4*(1+B) algorithmic bytes  =>  bytes/s allowed by the multiplier pipe for the layer's dominant height 2^15, B = 4.
        prod_per_s = 12.1 * N_SMS * sm_hz
        lde["product_pipe_bound_GBs"] = prod_per_s / (5 * 15 / 2 * 4 / 3) * 20 / 1e9
        lde["frac_of_product_pipe_bound"] = lde["achieved"] / lde["product_pipe_bound_GBs"]
        tree_h2d = {k: lanes[0].kind[k][2].h2d_bytes for k in lanes[0].kind}
        per_tree_h2d = sum(tree_h2d[kind_of_level(l)] * (args.leaves >> l) for l in range(depth + 1))
        per_tree_d2h = 4 * sum(sizes[l] * (args.leaves >> l) for l in range(depth + 1))
        line = {
            "metric": METRIC, "value": value, "unit": "aggregation proofs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_res / args.steps, "ms_per_layer": ms_layer,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 (31-bit Montgomery field, degree-4 extension)", "data": "synthetic",
            "config": {"workload": f"2-to-1 aggregation tree, {args.leaves} base proofs per tree ({args.field}): "
                                   f"{trees_per_step} trees per step (2 per GPU), {n_trees} trees in the timed region without a "
                                   f"barrier between them; nodes = steady-state recursion layer, level 1 = half-size layer with "
                                   f"packing (2,2), leaves = base Fibonacci circuit (D=1); FRI log_final_poly_len 6, 54 queries, "
                                   f"15-bit PoW; ms_per_layer: one node-shaped proof alone with the recursive_fibonacci FRI "
                                   f"parameters (log_final_poly_len 5)",
                       "shapes": {k: v.shapes for k, v in shapes.items()}, "fri_tree": fri_params(lib, 6), "fri_layer": lib.DEFAULT_FRI,
                       "scale": args.scale,
                       "l2": "part 1: flushed between timed steps (256 MiB fill); tree regions: no step boundary exists (continuous "
                             "pipeline) — the proofs in flight stream > 500 MB of LDE / digest data per GPU, 4x the 126 MB L2",
                       "parallelism": f"{world} GPU(s) x {args.inflight} proofs in flight; nodes of a tree on different ranks "
                                      f"(block partition rotated per tree); child proofs by NCCL send/recv, posting order = "
                                      + (f"wave (tree + {args.skew} x level)" if args.order == "wave" else
                                         f"blocks of {args.skew} trees, bucket = block + level, higher levels first inside a bucket"),
                       "proofs_per_step": trees_per_step * (2 * args.leaves - 1), "host_wait": wait_mode,
                       "high_priority_shapes": list(prio),
                       "host_cores": len(os.sched_getaffinity(0)), "proof_words": sizes, "layer_proof_words": proof_words,
                       "prep_commit_ms": prep_commit_ms, "launches_per_layer_proof": launches_per_proof},
            "e2e": {"value": e2e_value, "unit": "aggregation proofs/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": per_tree_h2d * trees_per_step, "d2h_bytes_per_step": per_tree_d2h * trees_per_step,
                    "wall_ms_per_step": wall_e2e / args.steps},
            "tree": {"leaves": args.leaves, "trees": n_trees, "aggregation_proofs": n_agg * n_trees,
                     "all_proofs": proved_total, "all_proofs_per_s": proved_total / (ms_res / 1e3),
                     "p2p_bytes_resident_region": sent_res, "p2p_bytes_e2e_region": sent_e2e,
                     "roots_checksum": rsum_res, "roots_checksum_e2e": rsum_e2e, "roots_match": rsum_res == rsum_e2e,
                     "roots_checksum_scope": f"weighted sum of the root-proof checksums of the first {common_trees} trees of the "
                                             "timed region (the trees every N proves): equal at every N iff every child proof "
                                             "reached its parent",
                     "lane_idle_fraction_worst_rank": idle_worst, "host_cores_busy_worst_rank": host_cores_busy,
                     "proof_ms_by_shape_rank0": r_res["proof_ms"], "node_proofs_in_flight_hist_rank0": r_res["nodes_in_flight"],
                     "per_rank": {"columns": ["ms_resident", "ms_e2e", "sm_mhz", "host_cores_busy", "lane_idle_fraction", "proofs"],
                                  "rows": per_rank},
                     "wall_ms_per_step": wall_res / args.steps, "lane_idle_fraction": idle_us / 1e6 / world / (wall_res / 1e3),
                     "critical_path_speedup_one_tree": agg.critical_path_speedup(args.leaves, world)},
            "gpu_launches": launches_res,
            "roofline": roofline,
            "roofline_lde": lde,
            "kernel_breakdown_ms": {k: round(v, 4) for k, v in breakdown.items()},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0))
            os.environ["OMP_NUM_THREADS"] = str(cores)
            dt, n, _ = time_oracle(args.field, fri_params(lib, 6), shapes["node"], 3, 1, args.cpu_baseline_budget_s)
            # the same inputs through a proving lane (an all-zero patch leaves the resident witness as generated): the
            # headline-size proof must be the oracle's, word for word — outside every timed region
            gpu_proof = lanes[0].prove("node", np.zeros(PATCH_WORDS, dtype=np.uint32), False)
            same = bool(np.array_equal(gpu_proof, time_oracle.last_proof))
            line["parity"] = {"check": "full-size aggregation-node proof: CUDA path == CPU oracle prover, all "
                                       f"{int(gpu_proof.size)} words", "ok": same}
            if not same:
                raise RuntimeError("bench: the GPU proof of the full-size node layer differs from the oracle's proof")
            line["cpu_baseline"] = {"value": 1.0 / dt, "unit": "proofs/s", "cores": cores, "kind": "port",
                                    "sample": f"{n} full-size aggregation-node layer proofs by oracle/liboracle.so (OpenMP, {cores} "
                                              f"threads), {dt:.2f} s each, marshalling outside the loop, no extrapolation; the "
                                              "leaf / level-1 proofs are not charged to the CPU",
                                    "published_reference": "109 ms/layer on an Apple M4 Pro, 14 cores (Rust prover, other hardware)"}
        print(json.dumps(line), file=OUT, flush=True)
    for ln in lanes:
        ln.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--field", default="koala-bear", choices=["koala-bear", "baby-bear"])
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--wait", choices=["spin", "yield", "block", "sleep"], default=None,
                    help="host wait mode of the prover threads (default: the library default, yield)")
    ap.add_argument("--inflight", type=int, default=4, help="concurrent proofs (lanes) per GPU")
    ap.add_argument("--leaves", type=int, default=8, help="base proofs per aggregation tree (power of two)")
    ap.add_argument("--trees-per-step", type=int, default=0, help="default 2 per GPU")
    ap.add_argument("--skew", type=int, default=8, help="wave skew of the hand-off posting order (aggregation.message_plan)")
    ap.add_argument("--tree-timeout-s", type=float, default=300.0)
    ap.add_argument("--clock-sample-ms", type=float, default=5.0, help="period of the NVML clock / throttle-reason samples")
    ap.add_argument("--comm-poll-us", type=float, default=200.0, help="idle period of the hand-off thread's poll")
    ap.add_argument("--priority-shapes", default="", help="comma list of tree shapes (leaf,l1,node) proved on high-priority streams")
    ap.add_argument("--order", default="block", choices=["wave", "block"], help="task / hand-off order (aggregation.order_key)")
    ap.add_argument("--pin-cores", type=int, default=0, help="1: give every rank its own slice of the host cores (sched_setaffinity)")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0, help="--impl reference: wall-clock budget of the whole run")
    ap.add_argument("--cpu-baseline-budget-s", type=float, default=30.0)
    args = ap.parse_args()
    # stdout carries exactly one JSON line: keep the real stdout aside and point fd 1 at stderr, so that anything a library
    # prints there (NCCL's version banner, for one) cannot precede it.
    global OUT
    sys.stdout.flush()
    OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

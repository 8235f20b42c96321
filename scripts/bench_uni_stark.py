"""BASELINE configs[2] base layer: one wide table (KeccakAir shape: 2 600 columns x 4 096 rows, recursive_keccak.rs:22-24,513-530)
proved in uni-stark mode on one GPU. Synthetic AIR of the same shape (airs/wide.py). Prints one JSON line: ms per proof alone
(device-resident trace and host matrix in), kernel-class breakdown. Usage: bench_uni_stark.py [width] [log_rows] [iters]"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = importlib.import_module("plonky3-recursion_b200.lib")
air = importlib.import_module("plonky3-recursion_b200.air")
wide = importlib.import_module("plonky3-recursion_b200.airs.wide")
fm = importlib.import_module("plonky3-recursion_b200.field")
width = int(sys.argv[1]) if len(sys.argv) > 1 else 2600
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
F = fm.get_field("koala-bear")
inst = air.build_instance("wide", wide.make_eval(width), F.p, log_n, width, 0, wide.N_PUBLIC, air.BusRegistry())
t, pubs = wide.trace(F.p, width, log_n)
ctx = lib.Context("koala-bear", lib.DEFAULT_FRI)
ctx.set_uni_stark(True)
pd = lib.ProverData.from_airs_and_degrees(ctx, [inst], [None])
prover = lib.BatchStarkProver(ctx, pinned_output=True)
tb_res = lib.TraceBatch(ctx, [t], [pubs], insts=[inst]).upload(pd)
tb_pin = lib.TraceBatch(ctx, [t], [pubs], pinned=True, insts=[inst])
for _ in range(3):
    prover.prove_resident(tb_res, pd, copy=False)
    prover.prove_all_tables(tb_pin, pd, copy=False)


def timed(fn):
    ms = []
    for _ in range(iters):
        ctx.timer_start()
        fn()
        ms.append(ctx.timer_stop())
    return float(np.median(ms))


ms_res = timed(lambda: prover.prove_resident(tb_res, pd, copy=False))
ms_host = timed(lambda: prover.prove_all_tables(tb_pin, pd, copy=False))
ctx.reset_kernel_stats()
ctx.set_kernel_timing(lib.KERNEL_CLASSES)
for _ in range(3):
    prover.prove_resident(tb_res, pd, copy=False)
st = ctx.kernel_stats()
print(json.dumps({"workload": f"uni-stark, one table {1 << log_n} x {width} (KeccakAir shape, synthetic constraints), koala-bear, "
                              "log_blowup 2, 54 queries, 15-bit PoW",
                  "ms_per_proof_resident": ms_res, "ms_per_proof_host_matrix_in": ms_host, "h2d_bytes": int(t.size * 4),
                  "proof_words": int(prover.last_proof_words), "constraints": int(inst.constraints.n_constraints),
                  "classes_ms": {k: round(v["ms"] / 3, 4) for k, v in st.items()},
                  "published_reference": "720 ms for the Keccak base proof (CPU p3-uni-stark, Apple M4 Pro; BASELINE.md) — other hardware, "
                                         "real KeccakAir constraints"}), flush=True)

"""A/B timing of one library build (P3R_LIB=... selects it): single-proof latency of the full-size layer (median of N, L2 not
flushed) and the per-kernel-class breakdown. Prints one JSON line. Usage: ab_time.py [field] [scale] [iters]"""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")
fm = importlib.import_module("plonky3-recursion_b200.field")
field = sys.argv[1] if len(sys.argv) > 1 else "koala-bear"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
F = fm.get_field(field)
L = wl.synthetic_layer(F, 1, n_const=int(1500 * scale), n_public=int(43000 * scale), n_alu=int(60000 * scale),
                       n_perms=int(12000 * scale), n_recompose=int(4000 * scale), min_height=256)
ctx = lib.Context(field)
spec = int(os.environ.get("P3R_SPEC", "1"))
ctx.lib.p3r_set_specialization(ctx.h, spec)
pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
prover = lib.BatchStarkProver(ctx, pinned_output=True)
tb = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
for _ in range(5):
    proof = prover.prove_resident(tb, pd)
ms = []
for _ in range(iters):
    ctx.timer_start()
    prover.prove_resident(tb, pd, copy=False)
    ms.append(ctx.timer_stop())
ctx.reset_kernel_stats()
ctx.set_kernel_timing(lib.KERNEL_CLASSES)
for _ in range(4):
    prover.prove_resident(tb, pd, copy=False)
st = ctx.kernel_stats()
print(json.dumps({"lib": os.path.basename(os.environ.get("P3R_LIB", "libp3r_b200.so")), "spec": spec,
                  "stage_max_log": os.environ.get("P3R_STAGE_MAX_LOG"), "field": field, "scale": scale,
                  "ms_median": float(np.median(ms)), "ms_min": float(min(ms)),
                  "digest": int(proof.astype(np.uint64).sum() % (1 << 61)),
                  "phases": {k: round(v, 3) for k, v in ctx.last_phase_times().items()},
                  "classes_ms": {k: round(v["ms"] / 4, 4) for k, v in st.items()},
                  "launches": {k: v["launches"] // 4 for k, v in st.items() if v["launches"]}}), flush=True)

"""Host CPU time per proof: wall clock vs CPU time of the calling thread (CLOCK_THREAD_CPUTIME_ID) around p3r_prove_resident,
with the driver's spin wait and with blocking waits (where CPU time = the host work: launches, transcript, staging)."""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")

ctx = lib.Context("koala-bear", lib.DEFAULT_FRI)
L = wl.synthetic_layer(ctx.field, 1, n_const=1500, n_public=43000, n_alu=60000, n_perms=12000, n_recompose=4000, min_height=256)
pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
prover = lib.BatchStarkProver(ctx, pinned_output=True)
tb = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
for mode in ("spin", "block", "yield"):
    ctx.set_wait_mode(mode)
    for _ in range(5):
        prover.prove_resident(tb, pd, copy=False)
    n = 50
    w0, c0, p0 = time.perf_counter(), time.thread_time(), time.process_time()
    for _ in range(n):
        prover.prove_resident(tb, pd, copy=False)
    w, c, p = time.perf_counter() - w0, time.thread_time() - c0, time.process_time() - p0
    print(f"{mode:6s} wall {w / n * 1e3:.3f} ms/proof   thread CPU {c / n * 1e3:.3f} ms/proof   process CPU {p / n * 1e3:.3f} ms/proof", flush=True)

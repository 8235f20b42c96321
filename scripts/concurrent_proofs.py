"""Throughput with several proofs in flight on ONE GPU (one context + stream + host thread per proof stream).
Usage: concurrent_proofs.py [n_streams] [steps]"""
import importlib, os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")
fm = importlib.import_module("plonky3-recursion_b200.field")
n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
F = fm.get_field("koala-bear")
L = wl.synthetic_layer(F, 1, n_const=1500, n_public=43000, n_alu=60000, n_perms=12000, n_recompose=4000, min_height=256)
workers = []
for k in range(n_streams):
    ctx = lib.Context("koala-bear")
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    prover = lib.BatchStarkProver(ctx, pinned_output=True)
    tb = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
    for _ in range(3):
        prover.prove_resident(tb, pd, copy=False)
    workers.append((ctx, pd, prover, tb))
for n in range(1, n_streams + 1):
    bar = threading.Barrier(n + 1)
    def run(w):
        ctx, pd, prover, tb = w
        bar.wait()
        for _ in range(steps):
            prover.prove_resident(tb, pd, copy=False)
        bar.wait()
    ths = [threading.Thread(target=run, args=(workers[k],)) for k in range(n)]
    for t in ths: t.start()
    bar.wait(); t0 = time.perf_counter(); bar.wait(); dt = time.perf_counter() - t0
    for t in ths: t.join()
    print(f"{n} proof stream(s): {n * steps / dt:.1f} proofs/s  ({dt / steps * 1e3:.2f} ms per round of {n})", flush=True)

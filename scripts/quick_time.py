"""Ad-hoc: time the full-size synthetic recursion layer on the GPU and print per-phase device times."""
import importlib, sys, time, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")
fm = importlib.import_module("plonky3-recursion_b200.field")
field = sys.argv[1] if len(sys.argv) > 1 else "koala-bear"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
F = fm.get_field(field)
t = time.time()
L = wl.synthetic_layer(F, 1, n_const=int(1500*scale), n_public=int(43000*scale), n_alu=int(60000*scale), n_perms=int(12000*scale), n_recompose=int(4000*scale), min_height=256)
print("gen %.1fs" % (time.time()-t), L.shapes, flush=True)
ctx = lib.Context(field)
t = time.time(); pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps); print("prep %.3fs" % (time.time()-t))
prover = lib.BatchStarkProver(ctx)
tb = lib.TraceBatch(ctx, L.traces, L.pubs)
for it in range(6):
    l0 = ctx.launch_count(); t = time.time(); proof = prover.prove_all_tables(tb, pd); dt = time.time()-t
    print("prove %.2f ms  launches %d  words %d " % (dt*1e3, ctx.launch_count()-l0, proof.size), {k: round(v,3) for k,v in ctx.last_phase_times().items()}, flush=True)
if len(sys.argv) > 3:
    from common import make_oracle
    orc = make_oracle(field, lib.DEFAULT_FRI)
    t=time.time(); orc.verify(L.insts, pd.preprocessed_commitment, L.pubs, proof); print("oracle verify ok %.2fs" % (time.time()-t))

"""Times p3r_prep_commit (ProverData::from_airs_and_degrees) on the full-size layer; with P3R_TRACE_PREP=1 the
library prints a per-stage breakdown (stream-synchronised laps) to stderr."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")

field = sys.argv[1] if len(sys.argv) > 1 else "koala-bear"
ctx = lib.Context(field, lib.DEFAULT_FRI)
L = wl.synthetic_layer(ctx.field, 1, n_const=1500, n_public=43000, n_alu=60000, n_perms=12000, n_recompose=4000, min_height=256)
for k in range(4):
    if k == 3:
        os.environ["P3R_TRACE_PREP"] = "1"
    pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
    print(f"prep_commit call {k}: {pd.commit_ms:.2f} ms", flush=True)
    pd.close()

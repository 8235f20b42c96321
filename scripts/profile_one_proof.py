"""ncu target: warm up, then bracket exactly ONE full-size layer proof with cudaProfilerStart/Stop
(use with `ncu --profile-from-start off ...`). Usage: profile_one_proof.py [field] [scale]"""
import ctypes, importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")
fm = importlib.import_module("plonky3-recursion_b200.field")
field = sys.argv[1] if len(sys.argv) > 1 else "koala-bear"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
F = fm.get_field(field)
L = wl.synthetic_layer(F, 1, n_const=int(1500 * scale), n_public=int(43000 * scale), n_alu=int(60000 * scale),
                       n_perms=int(12000 * scale), n_recompose=int(4000 * scale), min_height=256)
ctx = lib.Context(field)
pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
prover = lib.BatchStarkProver(ctx)
tb = lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)
for _ in range(3):
    prover.prove_resident(tb, pd)
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaProfilerStart()
l0 = ctx.launch_count()
prover.prove_resident(tb, pd)
rt.cudaProfilerStop()
print("profiled one proof:", ctx.launch_count() - l0, "launches", L.shapes)

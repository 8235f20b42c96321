"""Isolated sweep (BASELINE.json configs[4]): batched coset LDE + Poseidon2 Merkle commit on synthetic
matrices, one JSON line per shape. LDE: achieved ALGORITHMIC GB/s = 4*n*c*(1+B) bytes / time against the measured HBM peak;
Merkle: Poseidon2 permutations per ns. Usage: sweep_commit.py [field] [quick]"""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
lib = importlib.import_module("plonky3-recursion_b200.lib")
field = sys.argv[1] if len(sys.argv) > 1 else "koala-bear"
quick = len(sys.argv) > 2
peak = 6546.6
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = float(json.load(open(p))["hbm_gbs"])
shapes = [(15, 256), (18, 64), (18, 256), (20, 64), (20, 256)] + ([] if quick else [(22, 64), (22, 256)])
for log_blowup in (1, 2, 3):
    fri = dict(lib.DEFAULT_FRI)
    fri["log_blowup"] = log_blowup
    ctx = lib.Context(field, fri)
    for log_n, cols in shapes:
        if (log_n >= 22 and log_blowup == 3) or (log_n == 24 and log_blowup > 1):
            continue
        r = ctx.bench_commit(log_n, cols, iters=3)
        n, B = 1 << log_n, 1 << log_blowup
        lde_bytes = 4.0 * n * cols * (1 + B)
        perms = n * B * ((cols + 7) // 8) + n * B - 1
        gbs = lde_bytes / 1e9 / (r["lde_ms"] / 1e3)
        print(json.dumps({"field": field, "log_rows": log_n, "cols": cols, "blowup": B, "lde_ms": round(r["lde_ms"], 3),
                          "lde_algorithmic_GBps": round(gbs, 1), "lde_hbm_frac": round(gbs / peak, 4),
                          "merkle_ms": round(r["merkle_ms"], 3), "merkle_perms_per_ns": round(perms / (r["merkle_ms"] * 1e6), 3)}),
              flush=True)
    if log_blowup == 2:
        # FRI commit rounds: fold 2^(k + log_blowup) extension elements by arity 2/4/8 and commit the folded rows
        for log_n in (18, 20) if quick else (18, 20, 22):
            for log_arity in (1, 2, 3):
                log_len = log_n + log_blowup
                r = ctx.bench_fri_round(log_len, log_arity, iters=3)
                L = 1 << log_len
                fold_bytes = 16.0 * (L + (L >> log_arity))
                print(json.dumps({"field": field, "fri_log_len": log_len, "arity": 1 << log_arity, "fold_ms": round(r["fold_ms"], 4),
                                  "fold_GBps": round(fold_bytes / 1e9 / (r["fold_ms"] / 1e3), 1),
                                  "fold_hbm_frac": round(fold_bytes / 1e9 / (r["fold_ms"] / 1e3) / peak, 4),
                                  "commit_ms": round(r["commit_ms"], 3)}), flush=True)
    ctx.close()

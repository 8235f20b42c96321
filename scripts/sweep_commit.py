"""Isolated sweep (BASELINE.json configs[4]): batched coset LDE + Poseidon2 Merkle commit on synthetic
matrices — rows 2^18..2^24 x columns 64..512 at blowup 2 / 4 / 8 (every shape whose LDE has at most 2^32 elements: 16 GB), one
mixed-height batch {2^k, 2^(k-1), 2^(k-3)} per blowup, and FRI fold + commit rounds at arities 2 / 4 / 8 — one JSON line each.
LDE: achieved ALGORITHMIC GB/s = 4*n*c*(1+B) bytes / time against the measured HBM peak and against the integer-multiplier bound
(12.1 Montgomery products / clk / SM, (1+B)*log2(n)/2 butterflies per input element at ~4/3 products each); Merkle: Poseidon2
permutations per ns. A fresh context per shape (the arena keeps its slabs). Usage: sweep_commit.py [field] [quick]"""
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
PRODUCTS_PER_S = 12.1 * 148 * 1.965e9


def product_bound_gbs(log_n, B):
    return PRODUCTS_PER_S / ((1 + B) * log_n / 2 * 4 / 3) * 4 * (1 + B) / 1e9


rows = (18, 20) if quick else (18, 20, 22, 24)
TWO_ADICITY = {"koala-bear": 24, "baby-bear": 27}[field]   # 2^24-row LDEs exist only for BabyBear (KoalaBear: 2^24 is the largest domain)
for log_blowup in (1, 2, 3):
    B = 1 << log_blowup
    fri = dict(lib.DEFAULT_FRI)
    fri["log_blowup"] = log_blowup
    for log_n in rows:
        for cols in (64, 128, 256, 512):
            if (cols << (log_n + log_blowup)) > (1 << 32) or log_n + log_blowup > TWO_ADICITY:
                continue
            ctx = lib.Context(field, fri)
            r = ctx.bench_commit(log_n, cols, iters=2 if log_n >= 22 else 3)
            ctx.close()
            n = 1 << log_n
            lde_bytes = 4.0 * n * cols * (1 + B)
            perms = n * B * ((cols + 7) // 8) + n * B - 1
            gbs = lde_bytes / 1e9 / (r["lde_ms"] / 1e3)
            print(json.dumps({"field": field, "log_rows": log_n, "cols": cols, "blowup": B, "lde_ms": round(r["lde_ms"], 3),
                              "lde_algorithmic_GBps": round(gbs, 1), "lde_hbm_frac": round(gbs / peak, 4),
                              "lde_frac_of_product_pipe_bound": round(gbs / product_bound_gbs(log_n, B), 3),
                              "merkle_ms": round(r["merkle_ms"], 3), "merkle_perms_per_ns": round(perms / (r["merkle_ms"] * 1e6), 3)}),
                  flush=True)
    # mixed heights {2^k, 2^(k-1), 2^(k-3)}, 128 / 256 / 64 columns
    k = 20 if quick else 21
    lh, wd = [k, k - 1, k - 3], [128, 256, 64]
    ctx = lib.Context(field, fri)
    r = ctx.bench_commit_multi(lh, wd, iters=3)
    ctx.close()
    lde_bytes = sum(4.0 * (1 << h) * w * (1 + B) for h, w in zip(lh, wd))
    perms = sum((1 << h) * B * ((w + 7) // 8) for h, w in zip(lh, wd)) + (1 << k) * B - 1 + sum((1 << h) * B for h in lh[1:])
    gbs = lde_bytes / 1e9 / (r["lde_ms"] / 1e3)
    print(json.dumps({"field": field, "mixed_log_rows": lh, "cols": wd, "blowup": B, "lde_ms": round(r["lde_ms"], 3),
                      "lde_algorithmic_GBps": round(gbs, 1), "lde_hbm_frac": round(gbs / peak, 4), "merkle_ms": round(r["merkle_ms"], 3),
                      "merkle_perms_per_ns": round(perms / (r["merkle_ms"] * 1e6), 3)}), flush=True)
    if log_blowup == 2:
        # FRI commit rounds: fold 2^(k + log_blowup) extension elements by arity 2/4/8 and commit the folded rows
        ctx = lib.Context(field, fri)
        for log_n in (18, 20) if quick else (18, 20, 22, 24):
            for log_arity in (1, 2, 3):
                log_len = log_n + log_blowup
                if log_len > TWO_ADICITY:
                    continue
                r = ctx.bench_fri_round(log_len, log_arity, iters=3)
                L = 1 << log_len
                fold_bytes = 16.0 * (L + (L >> log_arity))
                print(json.dumps({"field": field, "fri_log_len": log_len, "arity": 1 << log_arity, "fold_ms": round(r["fold_ms"], 4),
                                  "fold_GBps": round(fold_bytes / 1e9 / (r["fold_ms"] / 1e3), 1),
                                  "fold_hbm_frac": round(fold_bytes / 1e9 / (r["fold_ms"] / 1e3) / peak, 4),
                                  "commit_ms": round(r["commit_ms"], 3)}), flush=True)
        ctx.close()

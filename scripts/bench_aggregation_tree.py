#!/usr/bin/env python
"""2-to-1 aggregation tree over N GPUs with REAL layer proofs and NCCL proof hand-off (BASELINE.json configs[3]).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \\
      scripts/bench_aggregation_tree.py [--leaves 8] [--trees 4] [--leaf-scale 0.25]

Every rank owns the leaves [g*L/G, (g+1)*L/G) (plonky3-recursion_b200/aggregation.py); a node is proved where its leftmost leaf
lives, so only right-child proofs move (dist.send / dist.recv of the ~400 KB proof blob over NCCL, GPU to GPU). Leaf proofs are
layer proofs at `leaf-scale`, aggregation nodes are full-size layer proofs; a node is proved only after both child proofs are
on its rank (in the real system they feed the host runner that builds the verifier-circuit witness). Prints one JSON line
with aggregation proofs/s for one tree (critical path limited: 7/3 at >= 4 GPUs) and for `trees` pipelined trees.
"""
import argparse, importlib, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

lib = importlib.import_module("plonky3-recursion_b200.lib")
wl = importlib.import_module("plonky3-recursion_b200.workload")
fm = importlib.import_module("plonky3-recursion_b200.field")
agg = importlib.import_module("plonky3-recursion_b200.aggregation")
FULL = dict(n_const=1500, n_public=43000, n_alu=60000, n_perms=12000, n_recompose=4000)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--leaves", type=int, default=8)
    ap.add_argument("--trees", type=int, default=4)
    ap.add_argument("--leaf-scale", type=float, default=0.25)
    ap.add_argument("--field", default="koala-bear")
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    F = fm.get_field(args.field)
    ctx = lib.Context(args.field, lib.DEFAULT_FRI, device=local)

    def setup(scale):
        L = wl.synthetic_layer(F, 1, min_height=256, **{k: max(8, int(v * scale)) for k, v in FULL.items()})
        pd = lib.ProverData.from_airs_and_degrees(ctx, L.insts, L.preps)
        return pd, lib.TraceBatch(ctx, L.traces, L.pubs).upload(pd)

    pd_node, tb_node = setup(1.0)
    pd_leaf, tb_leaf = setup(args.leaf_scale)
    prover = lib.BatchStarkProver(ctx)
    n_node = prover.prove_resident(tb_node, pd_node).size   # warm-up + proof sizes (identical on every rank)
    n_leaf = prover.prove_resident(tb_leaf, pd_leaf).size
    sizes = {0: n_leaf}
    counts = {"node": 0, "leaf": 0, "sent_bytes": 0}

    def prove_leaf(i):
        counts["leaf"] += 1
        return prover.prove_resident(tb_leaf, pd_leaf)

    def prove_node(nd, left, right):
        assert left.size == sizes.get(nd.level - 1, n_node) and right.size == left.size
        counts["node"] += 1
        return prover.prove_resident(tb_node, pd_node)

    def send(proof, dst, tag):
        counts["sent_bytes"] += proof.size * 4
        dist.send(torch.from_numpy(proof.view(np.int32)).to(dev), dst=dst)

    level_of_tag = {}

    def recv(src, tag):
        n = level_of_tag["n"]
        buf = torch.empty(n, dtype=torch.int32, device=dev)
        dist.recv(buf, src=src)
        return buf.cpu().numpy().view(np.uint32)

    def run_one_tree():
        # proof sizes: children of level-1 nodes are leaves, all others are node proofs
        mine = {}
        for lvl, row in enumerate(agg.schedule(args.leaves, world)):
            level_of_tag["n"] = n_leaf if lvl == 1 else n_node
            for nd, own, kids in row:
                if lvl == 0:
                    if own == rank:
                        mine[nd] = prove_leaf(nd.index)
                    continue
                inputs = []
                for child, cown in kids:
                    if cown == rank and own != rank:
                        send(mine[child], own, 0)
                    if own == rank:
                        inputs.append(mine[child] if cown == rank else recv(cown, 0))
                if own == rank:
                    mine[nd] = prove_node(nd, inputs[0], inputs[1])
        return mine

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    run_one_tree()  # warm-up (NCCL channels)
    barrier()
    t0 = time.perf_counter()
    run_one_tree()
    barrier()
    t_one = time.perf_counter() - t0
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.trees):   # back-to-back trees: ranks that finish their subtree start the next tree's leaves
        run_one_tree()
    barrier()
    t_many = time.perf_counter() - t0
    t = torch.tensor([t_one, t_many], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        n_agg = args.leaves - 1
        print(json.dumps({
            "metric": "aggregation proofs/s (2-to-1 tree, real layer proofs, NCCL proof hand-off)", "n_gpus": world,
            "leaves": args.leaves, "aggregation_proofs_per_tree": n_agg, "one_tree_ms": float(t[0]) * 1e3,
            "one_tree_agg_proofs_per_s": n_agg / float(t[0]),
            "pipelined_trees": args.trees, "pipelined_agg_proofs_per_s": args.trees * n_agg / float(t[1]),
            "critical_path_ideal_speedup": agg.critical_path_speedup(args.leaves, world),
            "proof_words": {"node": int(n_node), "leaf": int(n_leaf)}, "leaf_scale": args.leaf_scale,
            "rank0_counts": counts}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Host logic of the multi-GPU aggregation tree, exercised with world_size 2 over gloo on the CPU."""
import hashlib
import importlib
import os
import socket

import numpy as np
import pytest

agg = importlib.import_module("plonky3-recursion_b200.aggregation")


def test_schedule_partitions_subtrees():
    for world in (1, 2, 4, 8):
        levels = agg.schedule(8, world)
        assert [len(r) for r in levels] == [8, 4, 2, 1]
        assert levels[-1][0][1] == 0  # the root is proved on rank 0
        for row in levels:
            for nd, own, kids in row:
                assert 0 <= own < world
                if kids:
                    assert own == kids[0][1]  # parent lives with its left child: only the right child's proof moves
        leaf_owners = [own for _, own, _ in levels[0]]
        assert leaf_owners == sorted(leaf_owners) and len(set(leaf_owners)) == min(world, 8)
    # a single tree's aggregation proofs cannot scale past the 3-level critical path: 7/3
    assert agg.critical_path_speedup(8, 1) == 1.0
    assert abs(agg.critical_path_speedup(8, 2) - 7 / 4) < 1e-9
    assert abs(agg.critical_path_speedup(8, 4) - 7 / 3) < 1e-9 and abs(agg.critical_path_speedup(8, 8) - 7 / 3) < 1e-9


def _fake_leaf(i):
    return np.frombuffer(hashlib.sha256(f"leaf{i}".encode()).digest(), dtype=np.uint32).copy()


def _fake_node(nd, l, r):
    h = hashlib.sha256(l.tobytes() + r.tobytes() + f"{nd.level}:{nd.index}".encode()).digest()
    return np.frombuffer(h, dtype=np.uint32).copy()


def _serial_root(n_leaves):
    cur = [_fake_leaf(i) for i in range(n_leaves)]
    lvl = 0
    while len(cur) > 1:
        lvl += 1
        cur = [_fake_node(agg.Node(lvl, i), cur[2 * i], cur[2 * i + 1]) for i in range(len(cur) // 2)]
    return cur[0]


def _worker(rank, world, port, n_leaves, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def send(proof, dst, tag):
        dist.send(torch.from_numpy(proof.view(np.int32).copy()), dst=dst, tag=tag)

    def recv(src, tag):
        buf = torch.empty(8, dtype=torch.int32)
        dist.recv(buf, src=src, tag=tag)
        return buf.numpy().view(np.uint32).copy()

    mine = agg.run_tree(rank, world, n_leaves, _fake_leaf, _fake_node, send, recv)
    np.save(os.path.join(out_dir, f"rank{rank}.npy"),
            np.array([[nd.level, nd.index] for nd in mine], dtype=np.int64))
    root = agg.Node(n_leaves.bit_length() - 1, 0)
    if root in mine:
        np.save(os.path.join(out_dir, "root.npy"), mine[root])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_leaves", [8, 2])
def test_tree_over_two_ranks_gloo(tmp_path, n_leaves):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, n_leaves, str(tmp_path)), nprocs=2, join=True)
    root = np.load(tmp_path / "root.npy")
    assert np.array_equal(root, _serial_root(n_leaves))  # same result as the reference's serial loop order
    proved = [set(map(tuple, np.load(tmp_path / f"rank{r}.npy").tolist())) for r in range(2)]
    assert not (proved[0] & proved[1])
    assert len(proved[0] | proved[1]) == 2 * n_leaves - 1  # every leaf and every aggregation node proved exactly once
    assert (n_leaves.bit_length() - 1, 0) in proved[0]


def test_single_rank_tree_matches_serial():
    mine = agg.run_tree(0, 1, 8, _fake_leaf, _fake_node, None, None)
    assert len(mine) == 15 and np.array_equal(mine[agg.Node(3, 0)], _serial_root(8))


# ---- pipelined executor (bench.py --gpus N): many trees, several lanes per rank, non-blocking hand-off ------------------
PROOF_WORDS = {0: 24, 1: 40, 2: 56, 3: 56}   # level -> proof size (leaf / level 1 / upper levels differ, as on the GPU)


def _x_leaf(lane, tree, index):
    h = hashlib.sha256(f"leaf{tree}:{index}".encode()).digest()
    return np.resize(np.frombuffer(h, dtype=np.uint32), PROOF_WORDS[0]).copy()


def _x_node(lane, tree, nd, left, right):
    h = hashlib.sha256(left.tobytes() + right.tobytes() + f"{tree}:{nd.level}:{nd.index}".encode()).digest()
    return np.resize(np.frombuffer(h, dtype=np.uint32), PROOF_WORDS[nd.level]).copy()


def _x_serial_root(tree, n_leaves):
    cur = [_x_leaf(0, tree, i) for i in range(n_leaves)]
    lvl = 0
    while len(cur) > 1:
        lvl += 1
        cur = [_x_node(0, tree, agg.Node(lvl, i), cur[2 * i], cur[2 * i + 1]) for i in range(len(cur) // 2)]
    return cur[0]


def test_message_plan_is_balanced_and_ordered():
    for world in (2, 4, 8):
        msgs = agg.message_plan(8, world, 2 * world, skew=3)
        assert msgs == sorted(msgs)
        for key, t, lvl, ci, src, dst in msgs:
            assert src != dst and key[0] == t + 3 * lvl
            assert src == agg.rotated_owner(agg.Node(lvl - 1, ci), 8, world, t)
            assert dst == agg.rotated_owner(agg.Node(lvl, ci // 2), 8, world, t)
        # a proof's own children arrive in strictly earlier waves than the message that carries it onwards
        # (parent level + 1), which is what makes in-order posting deadlock-free
        load = {}
        for t in range(2 * world):
            for lvl in range(1, 4):
                for i in range(8 >> lvl):
                    r = agg.rotated_owner(agg.Node(lvl, i), 8, world, t)
                    load[r] = load.get(r, 0) + 1
        assert len(set(load.values())) == 1 and len(load) == world   # every rank proves the same number of nodes


@pytest.mark.parametrize("order", ["wave", "block"])
def test_order_key_puts_children_before_parents(order):
    """What in-order posting needs: the key of a proof (= of the message that carries it on) is larger than the keys of both
    child proofs, for every tree, level and skew."""
    for skew in (1, 3, 8):
        for t in range(20):
            for lvl in range(1, 4):
                for i in range(8 >> lvl):
                    k = agg.order_key(t, lvl, i, skew, order)
                    assert agg.order_key(t, lvl - 1, 2 * i, skew, order) < k
                    assert agg.order_key(t, lvl - 1, 2 * i + 1, skew, order) < k
    # block order: inside one bucket a rank sees the higher levels first
    ks = sorted(agg.order_key(t, lvl, 0, 8, "block") + (lvl,) for t in range(32) for lvl in range(4))
    in_bucket = [k[-1] for k in ks if k[0] == 4]
    assert in_bucket == sorted(in_bucket, reverse=True) and set(in_bucket) == {0, 1, 2, 3}
    with pytest.raises(ValueError):
        agg.order_key(0, 0, 0, 8, "nope")


@pytest.mark.parametrize("order", ["wave", "block"])
@pytest.mark.parametrize("lanes", [1, 4])
def test_executor_single_rank_many_trees(lanes, order):
    ex = agg.TreeExecutor(0, 1, 8, lanes, _x_leaf, _x_node, order=order, skew=2)
    out = ex.run(5)
    assert out["proved"] == {0: 40, 1: 20, 2: 10, 3: 5} and out["sent_bytes"] == 0
    for t in range(5):
        assert np.array_equal(out["roots"][t], _x_serial_root(t, 8))
    out = ex.run(2, first_tree=5)
    assert sorted(out["roots"]) == [5, 6] and np.array_equal(out["roots"][6], _x_serial_root(6, 8))


def test_executor_propagates_prover_errors():
    def bad_node(lane, tree, nd, left, right):
        raise ValueError("boom")
    with pytest.raises(ValueError):
        agg.TreeExecutor(0, 1, 4, 3, _x_leaf, bad_node).run(2)


def _x_leaf_jitter(lane, tree, index):
    import random
    import time
    time.sleep(random.Random(tree * 131 + index * 7 + 1).random() * 0.003)     # uneven proof times, different on every rank
    return _x_leaf(lane, tree, index)


def _x_node_jitter(lane, tree, nd, left, right):
    import random
    import time
    time.sleep(random.Random(tree * 977 + nd.level * 31 + nd.index).random() * 0.006)
    return _x_node(lane, tree, nd, left, right)


def _x_worker(rank, world, port, n_leaves, n_trees, lanes, out_dir, order="wave", skew=6, jitter=False):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tr = agg.TorchTransport(dist, torch, torch.device("cpu"), agg.MSG_HEADER_WORDS + max(PROOF_WORDS.values()), 8)
    ex = agg.TreeExecutor(rank, world, n_leaves, lanes, _x_leaf_jitter if jitter else _x_leaf, _x_node_jitter if jitter else _x_node,
                          tr, PROOF_WORDS, timeout_s=120, order=order, skew=skew)
    out = ex.run(n_trees)
    for t, proof in out["roots"].items():
        np.save(os.path.join(out_dir, f"root{t}.npy"), proof)
    np.save(os.path.join(out_dir, f"stats{rank}.npy"),
            np.array([out["sent_bytes"], out["recv_bytes"]] + [out["proved"][l] for l in sorted(out["proved"])], dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n_leaves,n_trees,lanes,order,skew",
                         [(2, 8, 6, 3, "wave", 6), (4, 8, 8, 2, "wave", 6), (2, 2, 5, 2, "wave", 6),
                          (2, 8, 7, 3, "block", 2), (4, 8, 9, 2, "block", 4)])
def test_executor_over_gloo_ranks(tmp_path, world, n_leaves, n_trees, lanes, order, skew):
    """Same roots as the serial loop of the reference (recursive_aggregation.rs:676-704) for every tree, every node proved
    exactly once, and real bytes on the wire in both directions (the rotation makes every rank send and receive)."""
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_x_worker, args=(world, port, n_leaves, n_trees, lanes, str(tmp_path), order, skew), nprocs=world, join=True)
    for t in range(n_trees):
        assert np.array_equal(np.load(tmp_path / f"root{t}.npy"), _x_serial_root(t, n_leaves))
    stats = np.array([np.load(tmp_path / f"stats{r}.npy") for r in range(world)])
    depth = n_leaves.bit_length() - 1
    assert [int(x) for x in stats[:, 2:].sum(axis=0)] == [n_trees * (n_leaves >> l) for l in range(depth + 1)]
    assert stats[:, 0].sum() == stats[:, 1].sum() > 0
    if n_trees >= world:
        assert (stats[:, 0] > 0).all() and (stats[:, 1] > 0).all()


@pytest.mark.parametrize("world,lanes,order,skew", [(4, 4, "block", 8), (4, 4, "wave", 3), (8, 2, "block", 8)])
def test_executor_with_uneven_proof_times(tmp_path, world, lanes, order, skew):
    """bench.py's configuration in miniature (block order of 8 / wave order; 8 ranks = one leaf per rank and tree, so leaf
    proofs travel too) with proof times that differ per task: the in-order posting of sends and receives must not deadlock
    when ranks drift apart, and every root is the serial one."""
    import torch.multiprocessing as mp
    n_leaves, n_trees = 8, 20
    with socket.socket() as s_:
        s_.bind(("127.0.0.1", 0))
        port = s_.getsockname()[1]
    mp.spawn(_x_worker, args=(world, port, n_leaves, n_trees, lanes, str(tmp_path), order, skew, True), nprocs=world, join=True)
    for t in range(n_trees):
        assert np.array_equal(np.load(tmp_path / f"root{t}.npy"), _x_serial_root(t, n_leaves))
    stats = np.array([np.load(tmp_path / f"stats{r}.npy") for r in range(world)])
    assert [int(x) for x in stats[:, 2:].sum(axis=0)] == [n_trees * (n_leaves >> l) for l in range(4)]

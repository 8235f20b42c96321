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

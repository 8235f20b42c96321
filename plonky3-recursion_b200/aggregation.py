"""2-to-1 aggregation tree sharded over GPUs.

The reference proves the tree serially (/root/reference recursion/examples/recursive_aggregation.rs:679-704: `for pair_idx`),
although leaves and the pairs of each level are independent (book/src/user_guide/aggregation.md "Tree aggregation").
Here rank g owns the subtree over leaves [g*L/G, (g+1)*L/G); a node is proved on the rank that owns its LEFTMOST leaf, so
at the levels where siblings live on different ranks the right child's proof is sent to the left child's rank
(point-to-point only: NCCL send/recv over NVLink on the GPU box, gloo in the CPU tests). No collective is involved.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class Node:
    level: int   # 0 = leaf
    index: int   # position within the level

    def children(self):
        return Node(self.level - 1, 2 * self.index), Node(self.level - 1, 2 * self.index + 1)


def owner(node: Node, n_leaves: int, world: int) -> int:
    """Rank that proves `node`: owner of its leftmost leaf under the block partition of the leaves."""
    assert n_leaves % world == 0 or world % n_leaves == 0 or world <= n_leaves
    leftmost = node.index << node.level
    per = max(1, n_leaves // world)
    return min(leftmost // per, world - 1)


def schedule(n_leaves: int, world: int):
    """Levels bottom-up; each level is a list of (node, owner_rank, [(child, child_owner), ...])."""
    assert n_leaves & (n_leaves - 1) == 0 and n_leaves >= 1
    depth = n_leaves.bit_length() - 1
    levels = []
    for lvl in range(depth + 1):
        row = []
        for i in range(n_leaves >> lvl):
            nd = Node(lvl, i)
            kids = [] if lvl == 0 else [(c, owner(c, n_leaves, world)) for c in nd.children()]
            row.append((nd, owner(nd, n_leaves, world), kids))
        levels.append(row)
    return levels


def critical_path_speedup(n_leaves: int, world: int) -> float:
    """Ideal speed-up of ONE tree's aggregation proofs (leaves excluded) when every proof costs the same:
    serial count / number of sequential rounds with `world` provers (8 leaves: 7 proofs, 3 rounds at >= 4 GPUs -> 2.33x)."""
    total, rounds = 0, 0
    for row in schedule(n_leaves, world)[1:]:
        total += len(row)
        per_rank = {}
        for _, r, _ in row:
            per_rank[r] = per_rank.get(r, 0) + 1
        rounds += max(per_rank.values())
    return total / rounds


def run_tree(rank: int, world: int, n_leaves: int, prove_leaf, prove_node, send, recv):
    """Execute the tree on this rank. prove_leaf(index) -> proof; prove_node(node, left_proof, right_proof) -> proof;
    send(proof, dst_rank, tag) / recv(src_rank, tag) -> proof move a proof between ranks. Returns {node: proof} for the nodes
    this rank proved (the root is in rank 0's dict)."""
    mine = {}
    tag = 0
    for lvl, row in enumerate(schedule(n_leaves, world)):
        for nd, own, kids in row:
            tag += 1
            if lvl == 0:
                if own == rank:
                    mine[nd] = prove_leaf(nd.index)
                continue
            inputs = []
            for k, (child, cown) in enumerate(kids):
                ctag = tag * 4 + k
                if cown == rank and own != rank:
                    send(mine[child], own, ctag)
                if own == rank:
                    inputs.append(mine[child] if cown == rank else recv(cown, ctag))
            if own == rank:
                mine[nd] = prove_node(nd, inputs[0], inputs[1])
    return mine

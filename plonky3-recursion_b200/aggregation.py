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


# ------------------------------------------------------------------------------------------------------------------------
# Pipelined executor: many trees in flight, several proofs in flight per rank, non-blocking proof hand-off.
# ------------------------------------------------------------------------------------------------------------------------
import heapq
import threading
import time
from collections import deque

import numpy as np

MSG_HEADER_WORDS = 4   # tree, child level, child index, payload words


def rotated_owner(node: Node, n_leaves: int, world: int, tree: int) -> int:
    """Owner of `node` in tree number `tree`: the block partition of `owner`, rotated by the tree index. One tree has
    n_leaves - 1 aggregation proofs for `world` ranks (rank 0 would prove three of the seven of an 8-leaf tree on 8 GPUs, half
    of the ranks none); over `world` consecutive trees every rank plays every role once, so the load is even."""
    return (owner(node, n_leaves, world) + tree) % world


def order_key(tree: int, level: int, index: int, skew: int, order: str = "wave"):
    """Priority of the proof of node (level, index) of tree `tree` — and of the message that carries it to its parent.
    "wave": tree + skew * (level + 1), older trees first inside a wave. "block": the trees are taken in blocks of `skew`;
    bucket = block + level + 1, and inside a bucket the HIGHER levels come first, then tree order. Both orders give a proof a
    larger key than the proofs (and messages) it depends on — children sit one bucket / `skew` waves earlier — which is all
    in-order posting needs. "block" keeps proofs of one shape together on a rank (root, level-2, level-1, leaves of eight
    trees) whatever the number of ranks: small latency-bound proofs run among themselves instead of beside full-size
    ones."""
    if order == "block":
        return (tree // skew + level + 1, -level, tree, index)
    if order != "wave":
        raise ValueError(f"unknown order {order!r}")
    return (tree + skew * (level + 1), tree, level, index)


def message_plan(n_leaves: int, world: int, n_trees: int, skew: int = 6, order: str = "wave"):
    """Every cross-rank child-proof transfer of `n_trees` trees as (wave, tree, parent level, child index, src, dst), in the
    ONE global order all ranks post their sends / receives in. wave = tree + skew * parent level: a proof depends only on
    messages of smaller waves (its own children's), which makes in-order posting deadlock-free even when a rank's transfers
    share one stream (NCCL point-to-point on the parent communicator), and the skew keeps about skew * depth trees in flight
    across ranks instead of serialising tree t + 1 behind tree t's root."""
    depth = n_leaves.bit_length() - 1
    msgs = []
    for t in range(n_trees):
        for lvl in range(1, depth + 1):
            for i in range(n_leaves >> lvl):
                nd = Node(lvl, i)
                own = rotated_owner(nd, n_leaves, world, t)
                for c in nd.children():
                    cown = rotated_owner(c, n_leaves, world, t)
                    if cown != own:
                        msgs.append((order_key(t, lvl - 1, c.index, skew, order), t, lvl, c.index, cown, own))
    msgs.sort()
    return msgs


class TreeExecutor:
    """Proves `n_trees` 2-to-1 aggregation trees on this rank's share of the nodes.

    prove_leaf(lane, tree, index) -> proof and prove_node(lane, tree, node, left, right) -> proof run on `n_lanes` worker
    threads (one proving context each); proofs are 1-D uint32 arrays. A node becomes ready the moment both child proofs are
    on this rank; ready tasks are taken in wave order (the order their results are needed in). Child proofs whose parent
    lives on another rank go through `transport` (isend / irecv / done / finish, driven by ONE communication thread in the
    global order of `message_plan`, never blocking the host). There is no barrier between trees.
    `sizes[level]` = proof words of a level-`level` proof (known from the warm-up; both ends of a transfer need it).
    """

    def __init__(self, rank: int, world: int, n_leaves: int, n_lanes: int, prove_leaf, prove_node, transport=None,
                 sizes=None, skew: int = 6, timeout_s: float = 600.0, comm_poll_s: float = 0.0002, order: str = "wave"):
        self.rank, self.world, self.n_leaves, self.n_lanes = rank, world, n_leaves, n_lanes
        self.depth = n_leaves.bit_length() - 1
        self.prove_leaf, self.prove_node, self.transport = prove_leaf, prove_node, transport
        self.sizes, self.skew, self.timeout_s, self.comm_poll_s = sizes or {}, skew, timeout_s, comm_poll_s
        self.order = order
        order_key(0, 0, 0, skew, order)
        if world > 1 and transport is None:
            raise ValueError("world > 1 needs a transport")

    def _key(self, tree, level, index):
        return order_key(tree, level, index, self.skew, self.order)

    def run(self, n_trees: int, first_tree: int = 0):
        """Returns {"roots": {tree: proof} (trees whose root this rank proved), "proved": {level: count}, "sent_bytes",
        "recv_bytes", "idle_s": per-lane seconds spent waiting for a ready task}."""
        rank, world, L = self.rank, self.world, self.n_leaves
        trees = range(first_tree, first_tree + n_trees)
        lock = threading.Condition()
        ready, arrived, outbox, roots = [], {}, {}, {}
        proved = {lvl: 0 for lvl in range(self.depth + 1)}
        state = {"left": 0, "abort": None, "sent": 0, "recv": 0}
        idle = [0.0] * self.n_lanes
        for t in trees:
            for lvl in range(self.depth + 1):
                for i in range(L >> lvl):
                    if rotated_owner(Node(lvl, i), L, world, t) == rank:
                        state["left"] += 1
                        if lvl == 0:
                            heapq.heappush(ready, (self._key(t, 0, i), t, 0, i))
        msgs = [m for m in message_plan(L, world, first_tree + n_trees, self.skew, self.order)
                if m[1] >= first_tree and rank in (m[4], m[5])] if world > 1 else []
        deadline = time.monotonic() + self.timeout_s

        def deliver(t, child: Node, proof):     # lock held
            parent = Node(child.level + 1, child.index // 2)
            slot = arrived.setdefault((t, parent), [None, None])
            slot[child.index & 1] = proof
            if slot[0] is not None and slot[1] is not None:
                heapq.heappush(ready, (self._key(t, parent.level, parent.index), t, parent.level, parent.index))
                lock.notify_all()

        def complete(t, nd: Node, proof):
            with lock:
                proved[nd.level] += 1
                state["left"] -= 1
                if nd.level == self.depth:
                    roots[t] = proof
                else:
                    parent = Node(nd.level + 1, nd.index // 2)
                    if rotated_owner(parent, L, world, t) == rank:
                        deliver(t, nd, proof)
                    else:
                        outbox[(t, nd.level, nd.index)] = proof
                lock.notify_all()

        def fail(e):
            with lock:
                if state["abort"] is None:
                    state["abort"] = e
                lock.notify_all()

        def lane_main(k):
            try:
                while True:
                    t0 = time.perf_counter()
                    with lock:
                        while not ready and state["left"] > 0 and state["abort"] is None:
                            if not lock.wait(timeout=0.25) and time.monotonic() > deadline:
                                raise TimeoutError(f"rank {rank}: tree executor timed out ({state['left']} proofs left)")
                        if state["abort"] is not None or not ready:
                            return
                        _, t, lvl, i = heapq.heappop(ready)
                        kids = arrived.pop((t, Node(lvl, i)), None) if lvl else None
                    idle[k] += time.perf_counter() - t0
                    nd = Node(lvl, i)
                    proof = self.prove_leaf(k, t, i) if lvl == 0 else self.prove_node(k, t, nd, kids[0], kids[1])
                    complete(t, nd, proof)
            except BaseException as e:   # noqa: BLE001 - propagate to run()
                fail(e)

        def comm_main():
            tr = self.transport
            try:
                pending, i = deque(), 0
                while i < len(msgs) or pending:
                    progressed = False
                    for _ in range(len(pending)):
                        h, m = pending[0]
                        if not tr.done(h):
                            break
                        pending.popleft()
                        got = tr.finish(h)
                        _, t, lvl, ci, src, dst = m
                        if dst == rank:
                            hdr, body = got[:MSG_HEADER_WORDS], got[MSG_HEADER_WORDS:]
                            if [int(x) for x in hdr] != [t, lvl - 1, ci, body.size]:
                                raise RuntimeError(f"rank {rank}: received {list(map(int, hdr))}, expected {(t, lvl - 1, ci, body.size)}")
                            with lock:
                                state["recv"] += got.size * 4
                                deliver(t, Node(lvl - 1, ci), body)
                        progressed = True
                    if i < len(msgs) and len(pending) < tr.max_outstanding:
                        _, t, lvl, ci, src, dst = msgs[i]
                        if src == rank:
                            with lock:
                                proof = outbox.pop((t, lvl - 1, ci), None)
                            if proof is not None:
                                buf = np.empty(MSG_HEADER_WORDS + proof.size, dtype=np.uint32)
                                buf[:MSG_HEADER_WORDS] = (t, lvl - 1, ci, proof.size)
                                buf[MSG_HEADER_WORDS:] = proof
                                pending.append((tr.isend(buf, dst), msgs[i]))
                                state["sent"] += buf.size * 4
                                i += 1
                                progressed = True
                        else:
                            pending.append((tr.irecv(src, MSG_HEADER_WORDS + self.sizes[lvl - 1]), msgs[i]))
                            i += 1
                            progressed = True
                    if not progressed:
                        if state["abort"] is not None:
                            return
                        if time.monotonic() > deadline:
                            raise TimeoutError(f"rank {rank}: proof hand-off timed out at message {i}/{len(msgs)}")
                        with lock:
                            lock.wait(timeout=self.comm_poll_s)
            except BaseException as e:   # noqa: BLE001
                fail(e)

        threads = [threading.Thread(target=lane_main, args=(k,), daemon=True) for k in range(self.n_lanes)]
        if msgs:
            threads.append(threading.Thread(target=comm_main, daemon=True))
        for th in threads:
            th.start()
        for th in threads:
            th.join()
        if state["abort"] is not None:
            raise state["abort"]
        return {"roots": roots, "proved": proved, "sent_bytes": state["sent"], "recv_bytes": state["recv"], "idle_s": idle}


class TorchTransport:
    """Proof hand-off over torch.distributed point-to-point (NCCL send/recv over NVLink on the GPU box: pinned host ->
    device staging -> ncclSend ... ncclRecv -> device staging -> pinned host, because the parent's witness generation needs
    the bytes on its host; gloo in the CPU tests). Used by one thread only; nothing here blocks the host."""

    def __init__(self, dist, torch, device, max_words: int, max_outstanding: int = 32):
        self.dist, self.torch, self.device, self.max_outstanding = dist, torch, device, max_outstanding
        self.dev = torch.empty((max_outstanding, max_words), dtype=torch.int32, device=device)
        pin = device.type == "cuda"
        self.host = torch.empty((max_outstanding, max_words), dtype=torch.int32, pin_memory=pin)
        self.free = list(range(max_outstanding))

    def isend(self, arr: np.ndarray, dst: int):
        k, n = self.free.pop(), arr.size
        h = self.host[k, :n]
        h.numpy()[:] = arr.view(np.int32)
        d = self.dev[k, :n]
        d.copy_(h, non_blocking=True)
        return (self.dist.isend(d, dst=dst), k, 0)

    def irecv(self, src: int, n_words: int):
        k = self.free.pop()
        return (self.dist.irecv(self.dev[k, :n_words], src=src), k, n_words)

    def done(self, h) -> bool:
        if self.device.type == "cuda":
            return h[0].is_completed()      # NCCL: the work's CUDA event — a poll, the host never blocks
        h[0].wait()                         # gloo completes a work only inside wait(): block on the oldest transfer, which
        return True                         # the global posting order makes safe (see message_plan)

    def finish(self, h):
        w, k, n = h
        if self.device.type == "cuda":
            w.wait()                        # stream dependency only: the work has completed (done() polled it)
        out = None
        if n:
            hv = self.host[k, :n]
            hv.copy_(self.dev[k, :n])          # device -> pinned host (synchronous)
            out = hv.numpy().view(np.uint32).copy()
        self.free.append(k)
        return out

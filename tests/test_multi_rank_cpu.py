"""CPU, world_size 2 over gloo: the multi-GPU plumbing (DD broadcast, shot sharding, histogram merge).
Each rank's sampler here is the oracle's Philox walk over its shot range; on GPUs it is k_sample."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.multi_gpu import broadcast_dd, merge_counts, shard
from qiskit_jku_provider_b200.engine import SNODE_DTYPE


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, total, seed, text, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle
    nodes = None
    if rank == 0:
        nodes = np.zeros(5, dtype=SNODE_DTYPE)
        nodes["c0"] = [0, 0, 1, 2, 3]
        nodes["c1"] = [0, 4, 4, 1, 2]
        nodes["q0"] = [0.0, 0.25, 0.5, 0.75, 1.0]
    got, root, nq, mm = broadcast_dd(dist, nodes, 4, n, np.arange(n, dtype=np.uint8), src=0)
    assert root == 4 and nq == n and list(mm) == list(range(n))
    assert got["q0"].tolist() == [0.0, 0.25, 0.5, 0.75, 1.0] and got["c1"].tolist() == [0, 4, 4, 1, 2]
    lo, hi = shard(total, rank, world)
    r = oracle.run_oracle(text, shots=hi - lo, seed=seed, extra=["--shot-offset", str(lo)]) if hi > lo else {"counts": {}}
    keys = np.array([int(k[::-1], 2) for k in r["counts"]], dtype=np.uint64)  # char i = qubit i -> bit i
    vals = np.array(list(r["counts"].values()), dtype=np.uint64)
    dense = merge_counts(dist, keys, vals, n)
    sparse = merge_counts(dist, keys, vals, 40)  # force the all-gather path
    if rank == 0:
        q.put((dense, sparse))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_counts_equal_single_rank(oracle_mod, world):
    circ = C.random_circuit(5, 30, 2)
    total, seed = 3001, 17
    full = oracle_mod.run_oracle(circ, shots=total, seed=seed)["counts"]
    expect = {int(k[::-1], 2): v for k, v in full.items()}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, circ.n, total, seed, circ.text(), q)) for r in range(world)]
    for p in procs:
        p.start()
    dense, sparse = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert dense == expect
    assert sparse == expect


def _owner(keys, world):
    return (keys * np.uint64(0x9E3779B97F4A7C15) >> np.uint64(40)) % np.uint64(world)


def _exchange_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qiskit_jku_provider_b200.multi_gpu import exchange_by_owner
    rng = np.random.default_rng(100 + rank)
    keys = np.unique(rng.integers(0, 5000, size=3000 + 500 * rank).astype(np.uint64))
    counts = rng.integers(1, 9, size=len(keys)).astype(np.int64)
    own = _owner(keys, world).astype(np.int64)
    order = np.argsort(own, kind="stable")          # what jkq_counts_partition does on the device
    per = np.bincount(own, minlength=world).tolist()
    got_k, got_c = exchange_by_owner(dist, torch.from_numpy(keys[order].view(np.int64).copy()), torch.from_numpy(counts[order].copy()), per)
    gk = got_k.numpy().view(np.uint64)
    assert (_owner(gk, world) == rank).all()       # only keys this rank owns arrive
    merged = {}
    for k, c in zip(gk.tolist(), got_c.tolist()):   # what jkq_counts_merge does on the device
        merged[k] = merged.get(k, 0) + c
    q.put((rank, dict(zip(keys.tolist(), counts.tolist())), merged))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sparse_exchange_delivers_every_pair_to_its_owner(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_exchange_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want, have = {}, {}
    for rank, local, merged in got:
        for k, c in local.items():
            want[k] = want.get(k, 0) + c
        assert not (set(merged) & set(have))        # owners are disjoint
        have.update(merged)
    assert have == want

"""Shot sharding over GPUs: one process per GPU, ``torch.distributed`` plumbing.

The canonical DD is built on rank 0 only (hash-consing does not shard).  The finished, level-ordered
sampling DD -- 16 bytes per node -- is broadcast (NCCL over NVLink on GPUs, gloo in the CPU tests),
every rank samples the contiguous shot range ``shard(total, rank, world)`` with the Philox counter =
GLOBAL shot index, and the count histograms are merged: dense all-reduce when the outcome space is
small, otherwise an all-gather of (key, count) pairs reduced on every rank.  Because the stream of a
shot depends only on (seed, global shot index), counts are identical for every world size.
"""

import numpy as np

DENSE_MAX_BITS = 20


def shard(total, rank, world):
    """Contiguous shot range [lo, hi) of ``rank``; sizes differ by at most one."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dev(dist):
    import torch
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_dd(dist, nodes, root, n_qubits, measure_map, src=0):
    """Broadcast (snode array, root, n, clbit map) from ``src``; returns them on every rank."""
    import torch
    dev = _dev(dist)
    rank = dist.get_rank()
    hdr = torch.zeros(3 + 64, dtype=torch.int64, device=dev)
    if rank == src:
        hdr[0], hdr[1], hdr[2] = len(nodes), root, n_qubits
        hdr[3:3 + n_qubits] = torch.as_tensor(np.asarray(measure_map, dtype=np.int64), device=dev)
    dist.broadcast(hdr, src=src)
    n_nodes, root, n_qubits = int(hdr[0]), int(hdr[1]), int(hdr[2])
    mm = hdr[3:3 + n_qubits].cpu().numpy().astype(np.uint8)
    from .engine import SNODE_DTYPE
    buf = torch.zeros(n_nodes * 2, dtype=torch.int64, device=dev)  # 16 bytes per node
    if rank == src:
        buf.copy_(torch.from_numpy(np.ascontiguousarray(nodes).view(np.int64).copy()).to(dev))
    dist.broadcast(buf, src=src)
    out = buf.cpu().numpy().view(SNODE_DTYPE)
    return out, root, n_qubits, mm


def merge_counts(dist, keys, counts, n_bits):
    """All-reduce per-rank (keys, counts) into the global histogram, returned as a dict on every rank."""
    import torch
    dev = _dev(dist)
    keys = np.asarray(keys, dtype=np.uint64)
    counts = np.asarray(counts, dtype=np.uint64)
    if n_bits <= DENSE_MAX_BITS:
        dense = torch.zeros(1 << n_bits, dtype=torch.int64, device=dev)
        if len(keys):
            dense.index_add_(0, torch.as_tensor(keys.astype(np.int64), device=dev),
                             torch.as_tensor(counts.astype(np.int64), device=dev))
        dist.all_reduce(dense)
        nz = torch.nonzero(dense).flatten()
        return {int(k): int(v) for k, v in zip(nz.cpu().tolist(), dense[nz].cpu().tolist())}
    world = dist.get_world_size()
    sizes = torch.zeros(world, dtype=torch.int64, device=dev)
    sizes[dist.get_rank()] = len(keys)
    dist.all_reduce(sizes)
    cap = int(sizes.max())
    mine = torch.zeros(2 * cap, dtype=torch.int64, device=dev)
    if len(keys):
        mine[:len(keys)] = torch.as_tensor(keys.view(np.int64), device=dev)
        mine[cap:cap + len(keys)] = torch.as_tensor(counts.view(np.int64), device=dev)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    merged = {}
    for r, g in enumerate(gathered):
        g = g.cpu().numpy()
        k = g[:int(sizes[r])].view(np.uint64)
        v = g[cap:cap + int(sizes[r])]
        for a, b in zip(k.tolist(), v.tolist()):
            merged[a] = merged.get(a, 0) + b
    return merged


def sample_sharded(dist, engine, total_shots, seed, circuit=None, measure=None):
    """Run ``circuit`` on rank 0, broadcast its sampling DD, sample ``total_shots`` split over the
    ranks, return (merged counts dict, local engine stats).  ``engine``: this rank's Engine."""
    rank, world = dist.get_rank(), dist.get_world_size()
    nodes = root = n = mm = None
    if rank == 0:
        engine.load(circuit, measure=measure)
        engine.apply()
        nodes, root = engine.dd_export()
        n, mm = engine.n, engine.measure_map()
    if world > 1:
        nodes, root, n, mm = broadcast_dd(dist, nodes, root, n, mm, src=0)
        if rank != 0:
            engine.dd_import(n, nodes, root, mm)
    lo, hi = shard(total_shots, rank, world)
    engine.sample(hi - lo, seed=seed, shot_offset=lo)
    keys, vals = engine.counts_arrays()
    n_bits = int(sum(1 for b in mm if b < 64))
    if world > 1:
        return merge_counts(dist, keys, vals, n_bits), engine.stats()
    return {int(k): int(v) for k, v in zip(keys, vals)}, engine.stats()

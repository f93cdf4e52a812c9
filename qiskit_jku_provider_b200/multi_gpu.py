"""Shot sharding over GPUs: one process per GPU, ``torch.distributed`` plumbing.

The canonical DD is built on rank 0 only (hash-consing does not shard).  The finished, level-ordered
sampling DD -- 16 bytes per node -- is broadcast (NCCL over NVLink on GPUs, gloo in the CPU tests),
every rank samples the contiguous shot range ``shard(total, rank, world)`` with the Philox counter =
GLOBAL shot index, and the count histograms are merged: dense all-reduce when the outcome space is
small; otherwise the sparse exchange -- every key has an owner rank, each rank sorts its (key, count)
pairs by owner on the device (``jkq_counts_partition``), one all-to-all moves them (NCCL over NVLink), the
owner adds them into its histogram (``jkq_counts_merge``) and the result stays distributed by owner
(``merge_counts_by_owner``).  ``merge_counts`` (all-gather, dict on every rank) remains for small histograms.
Because the stream of a shot depends only on (seed, global shot index), counts are identical for every
world size.
"""

import numpy as np

DENSE_MAX_BITS = 20
SMALL_HISTOGRAM = 1 << 16  # up to this many distinct outcomes per rank the (key, count) pairs are simply all-gathered


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


INLINE_DD_NODES = 4096  # a sampling DD of up to this many nodes travels in the same broadcast as its header


def broadcast_dd_device(dist, engine, src=0):
    """NCCL path: the sampling DD travels device to device -- ``jkq_dd_export_device`` into a torch tensor on rank
    ``src``, NCCL broadcast over NVLink, ``jkq_dd_import_device`` everywhere else.  No host staging of the node
    array (the 3 + 64 header words are the only thing a rank reads back).  A DD of at most INLINE_DD_NODES nodes (GHZ:
    47, QFT-32: 32) rides in the header's broadcast: one collective.  Returns (n_qubits, measure map)."""
    import torch
    dev = _dev(dist)
    rank = dist.get_rank()
    H = 3 + 64
    buf = torch.zeros(H + 2 * INLINE_DD_NODES, dtype=torch.int64, device=dev)
    n_nodes = root = 0
    if rank == src:
        n_nodes, root = engine.dd_size()
        mm = engine.measure_map()
        h = np.zeros(H, dtype=np.int64)
        h[0], h[1], h[2] = n_nodes, root, engine.n
        h[3:3 + engine.n] = mm
        buf[:H].copy_(torch.from_numpy(h))
        if n_nodes <= INLINE_DD_NODES:
            engine.dd_export_device(buf.data_ptr() + 8 * H, n_nodes)
    dist.broadcast(buf, src=src)
    hh = buf[:H].cpu().numpy()
    n_nodes, root, n_qubits = int(hh[0]), int(hh[1]), int(hh[2])
    mm = hh[3:3 + n_qubits].astype(np.uint8)
    if n_nodes <= INLINE_DD_NODES:
        if rank != src:  # (.cpu() above synchronised the stream the broadcast ran on)
            engine.dd_import_device(n_qubits, buf.data_ptr() + 8 * H, n_nodes, root, mm)
        return n_qubits, mm
    big = torch.empty(n_nodes * 2, dtype=torch.int64, device=dev)  # 16 bytes per node
    if rank == src:
        engine.dd_export_device(big.data_ptr(), n_nodes)
    dist.broadcast(big, src=src)
    if rank != src:
        torch.cuda.current_stream().synchronize()  # the library copies on its own stream
        engine.dd_import_device(n_qubits, big.data_ptr(), n_nodes, root, mm)
    return n_qubits, mm


TINY_HISTOGRAM = 1024  # up to this many distinct outcomes per rank: one all-gather of fixed-size blocks merges them


def merge_counts_tiny(dist, engine):
    """A handful of outcomes per rank (GHZ: 2): every rank puts (n, keys, counts) into a fixed block, ONE all-gather,
    merged on the host.  Returns the dict, or None when some rank holds more than TINY_HISTOGRAM outcomes."""
    import torch
    dev = _dev(dist)
    world = dist.get_world_size()
    K = TINY_HISTOGRAM
    n = engine.distinct_outcomes()
    blk = np.zeros(1 + 2 * K, dtype=np.int64)
    if n <= K:
        keys, vals = engine.counts_arrays()
        blk[0] = len(keys)
        blk[1:1 + len(keys)] = keys.view(np.int64)
        blk[1 + K:1 + K + len(keys)] = vals.view(np.int64)
    else:
        blk[0] = -1
    mine = torch.from_numpy(blk).to(dev)
    allb = torch.empty(world * (1 + 2 * K), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allb, mine)
    a = allb.cpu().numpy().reshape(world, 1 + 2 * K)
    if (a[:, 0] < 0).any():
        return None
    merged = {}
    for r in range(world):
        m = int(a[r, 0])
        for k, v in zip(a[r, 1:1 + m].view(np.uint64).tolist(), a[r, 1 + K:1 + K + m].tolist()):
            merged[k] = merged.get(k, 0) + v
    return merged


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


def exchange_by_owner(dist, keys, counts, per_owner):
    """The one exchange step of the sparse merge: ``keys`` / ``counts`` (int64 tensors on the backend's device, sorted
    by owner rank, ``per_owner[r]`` entries for rank r) go out in an all-to-all; returns what this rank owns."""
    import torch
    dev = keys.device
    send_sizes = torch.tensor([int(x) for x in per_owner], dtype=torch.int64, device=dev)
    recv_sizes = torch.empty_like(send_sizes)
    dist.all_to_all_single(recv_sizes, send_sizes)
    rs = [int(x) for x in recv_sizes.tolist()]
    ss = [int(x) for x in per_owner]
    got_k = torch.empty(sum(rs), dtype=torch.int64, device=dev)
    got_c = torch.empty(sum(rs), dtype=torch.int64, device=dev)
    dist.all_to_all_single(got_k, keys, output_split_sizes=rs, input_split_sizes=ss)
    dist.all_to_all_single(got_c, counts, output_split_sizes=rs, input_split_sizes=ss)
    return got_k, got_c


def merge_counts_by_owner(dist, engine):
    """Sparse histogram merge on the GPUs: partition by owner -> all-to-all -> merge into the owner's histogram.
    Afterwards ``engine.counts_arrays()`` holds the outcomes THIS rank owns with their global counts.  Returns
    (distinct outcomes owned here, distinct outcomes in total, pairs sent, pairs received)."""
    import torch
    world = dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    total = sum(engine.counts_partition(world))
    keys = torch.empty(max(1, total), dtype=torch.int64, device=dev)
    counts = torch.empty(max(1, total), dtype=torch.int64, device=dev)
    per = engine.counts_partition(world, keys.data_ptr(), counts.data_ptr(), total)
    got_k, got_c = exchange_by_owner(dist, keys[:total], counts[:total], per)
    torch.cuda.synchronize()  # the library merges on its own stream
    engine.counts_merge(got_k.data_ptr() if got_k.numel() else 0, got_c.data_ptr() if got_c.numel() else 0, got_k.numel(),
                        expect=got_k.numel(), reset=True)
    owned = engine.distinct_outcomes()
    t = torch.tensor([owned], dtype=torch.int64, device=dev)
    dist.all_reduce(t)
    return owned, int(t[0]), total, int(got_k.numel())


def sample_sharded(dist, engine, total_shots, seed, circuit=None, measure=None, sparse="gather"):
    """Run ``circuit`` on rank 0, broadcast its sampling DD, sample ``total_shots`` split over the
    ranks, return (merged counts dict, local engine stats).  ``engine``: this rank's Engine.
    ``sparse="owner"`` (NCCL only): an outcome space too large for the dense all-reduce is merged by the
    owner exchange and stays distributed -- the first return value is then
    {"owned": distinct outcomes in this rank's histogram, "distinct_total": ..., "pairs_sent": ...}."""
    rank, world = dist.get_rank(), dist.get_world_size()
    nodes = root = n = mm = None
    if rank == 0:
        engine.load(circuit, measure=measure)
        engine.apply()
        n, mm = engine.n, engine.measure_map()
    if world > 1 and dist.get_backend() == "nccl":
        n, mm = broadcast_dd_device(dist, engine, src=0)  # device to device over NVLink
    elif world > 1:
        if rank == 0:
            nodes, root = engine.dd_export()
        nodes, root, n, mm = broadcast_dd(dist, nodes, root, n, mm, src=0)
        if rank != 0:
            engine.dd_import(n, nodes, root, mm)
    lo, hi = shard(total_shots, rank, world)
    engine.sample(hi - lo, seed=seed, shot_offset=lo)
    # histogram keys are indexed by classical-bit POSITION: the width is the highest position + 1, not the number of
    # measured qubits (a sparse map such as {0: 5} has one measured qubit and keys up to 32)
    n_bits = max([int(b) for b in mm if b < 64] + [0]) + 1
    if world > 1 and dist.get_backend() == "nccl":
        tiny = merge_counts_tiny(dist, engine)
        if tiny is not None:
            return tiny, engine.stats()
    if world > 1 and sparse == "owner" and n_bits > DENSE_MAX_BITS and dist.get_backend() == "nccl":
        # how large is the largest local histogram?  (counted by the sampling kernel itself: no pass over the table)
        import torch
        d = torch.tensor([engine.distinct_outcomes()], dtype=torch.int64, device=_dev(dist))
        dist.all_reduce(d, op=dist.ReduceOp.MAX)
        if int(d[0]) > SMALL_HISTOGRAM:
            owned, total, sent, _ = merge_counts_by_owner(dist, engine)
            return {"owned": owned, "distinct_total": total, "pairs_sent": sent}, engine.stats()
        n_bits = 64  # a handful of outcomes of a wide register (GHZ): the all-gather merge below
    keys, vals = engine.counts_arrays()
    if world > 1:
        return merge_counts(dist, keys, vals, n_bits), engine.stats()
    return {int(k): int(v) for k, v in zip(keys, vals)}, engine.stats()


def run_shots_sharded(dist, runner, n_qubits, ops, total_shots, seed, measure_seed0=None, physical_reset=False, n_bits=None):
    """Per-shot re-simulation of a measured circuit (``BatchRunner.run_shots``) with the shots split over the ranks:
    rank r runs the contiguous range ``shard(total, r, world)``; shot s uses the measurement stream
    ``measure_seed0 + s`` and read-out shot s of stream ``seed`` on whichever rank it lands, so the merged histogram
    does not depend on the world size.  Returns (counts dict on every rank, shots this rank could not finish on
    the single-CTA path, kernel ms of this rank)."""
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    if measure_seed0 is None:
        measure_seed0 = (int(seed) * 0x9E3779B97F4A7C15) & 0x3FFFFFFFFFFFFFFF
    lo, hi = shard(total_shots, rank, world)
    keys, _, status, _, ms = runner.run_shots(n_qubits, ops, hi - lo, seed=seed, shot_offset=lo, measure_seed0=measure_seed0 + lo,
                                              physical_reset=physical_reset)
    good = np.asarray(keys)[np.asarray(status) == 0]
    uniq, cnt = np.unique(good, return_counts=True) if len(good) else (np.zeros(0, np.uint64), np.zeros(0, np.int64))
    failed = [lo + int(i) for i in np.nonzero(np.asarray(status))[0]]
    if world > 1:
        if n_bits is None:  # widest classical bit any MEASURE / MEASURE_ONE op writes (+ 1), not the qubit count
            n_bits = max([int(o["arg"]) + 1 for o in ops if int(o["kind"]) in (2, 4)] + [1]) if hasattr(ops, "__iter__") and len(ops) and isinstance(ops[0], dict) else 64
        if len(uniq) and int(uniq.max()) >= (1 << min(int(n_bits), 63)):
            n_bits = 64  # a key outside the declared width: take the sparse merge
        return merge_counts(dist, uniq.astype(np.uint64), cnt.astype(np.uint64), int(n_bits)), failed, ms
    return {int(k): int(v) for k, v in zip(uniq, cnt)}, failed, ms

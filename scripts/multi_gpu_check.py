"""Run under torchrun with N >= 2 ranks (one per GPU): sharded sampling, the sparse owner exchange and sharded per-shot
re-simulation must give exactly what one GPU gives.  Prints MULTI_GPU_CHECK OK on rank 0."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
os.environ.setdefault("NCCL_DEBUG", "WARN")
from qiskit_jku_provider_b200 import circuits as C, multi_gpu  # noqa: E402
from qiskit_jku_provider_b200.batch import BatchRunner  # noqa: E402
from qiskit_jku_provider_b200.engine import Engine  # noqa: E402
from qiskit_jku_provider_b200.qasm import parse_qasm  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = Engine(device=local, node_capacity=1 << 16)
    single = Engine(device=local, node_capacity=1 << 16)
    report = {}
    # (a) dense merge: GHZ / random circuit on few qubits
    circ = C.random_circuit(10, 60, 3)
    counts, _ = multi_gpu.sample_sharded(dist, eng, 200001, 7, circuit=circ)
    single.load(circ)
    single.run(200001, 7)
    assert counts == single.counts(), "dense merge differs from one GPU"
    # (b) sparse exchange by owner: QFT-28, 2e6 shots -> ~2e6 distinct outcomes
    circ = C.qft(28)
    shots = 2000000
    nodes = root = n = mm = None
    if rank == 0:
        eng.load(circ)
        eng.apply()
        nodes, root = eng.dd_export()
        n, mm = eng.n, eng.measure_map()
    nodes, root, n, mm = multi_gpu.broadcast_dd(dist, nodes, root, n, mm, src=0)
    if rank != 0:
        eng.dd_import(n, nodes, root, mm)
    lo, hi = multi_gpu.shard(shots, rank, world)
    for attempt in range(2):  # (the first exchange also pays for NCCL's lazy p2p set-up)
        eng.sample(hi - lo, seed=5, shot_offset=lo)
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        owned, total, sent, received = multi_gpu.merge_counts_by_owner(dist, eng)
        torch.cuda.synchronize()
        dist.barrier()
        report["sparse_exchange_s_%s" % ("first" if attempt == 0 else "warm")] = time.perf_counter() - t0
    report["distinct_total"] = total
    report["pairs_sent_rank0"] = sent
    k, c = eng.counts_arrays()
    single.load(circ)
    single.run(shots, 5)
    sk, sc = single.counts_arrays()
    assert total == len(sk), "distinct outcomes: %d vs %d on one GPU" % (total, len(sk))
    want = dict(zip(sk.tolist(), sc.tolist()))
    assert all(want.get(a) == b for a, b in zip(k.tolist(), c.tolist())), "owned counts differ from one GPU"
    tot = torch.tensor([int(c.sum())], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    assert int(tot[0]) == shots
    # (c) per-shot re-simulation, shots split over the ranks
    from test_qasm_frontend import load_case
    prog = parse_qasm(load_case("teleport")[0])
    with BatchRunner(workers=1 << 16, device=local) as br:
        merged, failed, ms = multi_gpu.run_shots_sharded(dist, br, prog.n_qubits, prog.ops, 30001, 9)
        assert not failed
        if rank == 0:
            keys, _, status, _, _ = br.run_shots(prog.n_qubits, prog.ops, 30001, seed=9)
            u, cnt = np.unique(keys, return_counts=True)
            assert merged == dict(zip(u.tolist(), cnt.tolist())), "sharded per-shot counts differ from one GPU"
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_CHECK OK world=%d %s" % (world, report))
    eng.close()
    single.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""GPU: the device side of the sparse histogram exchange (jkq_counts_partition / jkq_counts_merge) on one
GPU -- partition a histogram for a virtual world of 4 ranks, check every pair lands in exactly one owner's range, and
merge the parts back -- and, when the box has two GPUs, the real thing under torchrun (scripts/multi_gpu_check.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_and_merge_round_trip():
    import torch
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import Engine
    with Engine(device=0, node_capacity=1 << 16) as eng:
        eng.load(C.qft(26))
        eng.run(300000, 11)            # ~300 000 distinct outcomes out of 2^26
        keys0, cnt0 = eng.counts_arrays()
        want = dict(zip(keys0.tolist(), cnt0.tolist()))
        assert len(want) > 290000 and sum(want.values()) == 300000
        for world in (1, 4, 7):
            per = eng.counts_partition(world)
            assert sum(per) == len(want) and (world == 1 or min(per) > 0.8 * len(want) / world)
            dev = torch.device("cuda", 0)
            k = torch.empty(len(want), dtype=torch.int64, device=dev)
            c = torch.empty(len(want), dtype=torch.int64, device=dev)
            assert eng.counts_partition(world, k.data_ptr(), c.data_ptr(), len(want)) == per
            kk, cc = k.cpu().numpy().view(np.uint64), c.cpu().numpy()
            assert dict(zip(kk.tolist(), cc.tolist())) == want          # every pair exactly once
            # owner ranges are consistent: the same key always falls into the same range for a given world
            bounds = np.cumsum([0] + per)
            with Engine(device=0, node_capacity=1 << 12) as other:      # a second handle plays "the owner"
                got = {}
                for r in range(world):
                    lo, hi = int(bounds[r]), int(bounds[r + 1])
                    other.counts_merge(k[lo:hi].data_ptr() if hi > lo else 0, c[lo:hi].data_ptr() if hi > lo else 0, hi - lo, expect=hi - lo, reset=True)
                    ok, oc = other.counts_arrays()
                    assert len(ok) == hi - lo
                    assert not (set(ok.tolist()) & set(got))
                    got.update(zip(ok.tolist(), oc.tolist()))
                assert got == want
        # merging a histogram into itself twice doubles the counts (accumulating merge)
        k = torch.from_numpy(keys0.view(np.int64).copy()).cuda()
        c = torch.from_numpy(cnt0.view(np.int64).copy()).cuda()
        torch.cuda.synchronize()
        eng.counts_merge(k.data_ptr(), c.data_ptr(), len(keys0), expect=len(keys0), reset=True)
        eng.counts_merge(k.data_ptr(), c.data_ptr(), len(keys0), reset=False)
        k2, c2 = eng.counts_arrays()
        assert dict(zip(k2.tolist(), c2.tolist())) == {a: 2 * b for a, b in want.items()}


def test_partition_needs_a_histogram_and_room():
    import torch
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import Engine
    from qiskit_jku_provider_b200.jkusimulatorerror import JKUSimulatorError
    with Engine(device=0, node_capacity=1 << 12) as eng:
        with pytest.raises(JKUSimulatorError):
            eng.counts_partition(2)
        eng.load(C.qft(10))
        eng.run(5000, 1)
        k = torch.empty(8, dtype=torch.int64, device="cuda")
        with pytest.raises(JKUSimulatorError):
            eng.counts_partition(2, k.data_ptr(), k.data_ptr(), 8)


def test_two_gpus_under_torchrun():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29533", os.path.join(ROOT, "scripts", "multi_gpu_check.py")],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0 and "MULTI_GPU_CHECK OK" in out.stdout, out.stdout[-3000:]


def test_sample_multi_gpu_single_process_equals_one_gpu():
    """jkq_sample_multi_gpu (C ABI): the sampling DD goes to helper engines by peer copy, the shots are split, the
    histograms merged on the device.  Here both "devices" are GPU 0 (two engines side by side), which runs the whole
    path on a one-GPU box; scripts/multi_gpu_check.py repeats it over real peers."""
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import Engine
    for circ, shots in ((C.ghz(24), 1 << 20), (C.qft(12), 300000), (C.random_circuit(10, 40, 3), 100001)):
        with Engine(device=0, node_capacity=1 << 18) as a, Engine(device=0, node_capacity=1 << 18) as b:
            a.load(circ)
            a.run(shots, 9)
            want = a.counts()
            b.load(circ)
            b.apply()
            for devs in ([0], [0, 0], [0, 0, 0]):
                st = b.sample_multi(devs, shots, 9)
                assert st["n_devices"] == len(devs) and st["merged"] == 1
                assert b.counts() == want, (circ.name, devs)
                assert st["distinct_total"] == len(want)


def test_dd_broadcast_through_device_buffers():
    """jkq_dd_export_device / jkq_dd_import_device: what the NCCL broadcast uses on both ends."""
    import torch
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import Engine
    circ = C.random_circuit(9, 50, 5)
    with Engine(device=0, node_capacity=1 << 16) as a, Engine(device=0, node_capacity=1 << 16) as b:
        a.load(circ)
        a.run(50000, 3)
        n_nodes, root = a.dd_size()
        buf = torch.empty(2 * n_nodes, dtype=torch.int64, device="cuda:0")
        assert a.dd_export_device(buf.data_ptr(), n_nodes) == (n_nodes, root)
        torch.cuda.synchronize()
        b.dd_import_device(circ.n, buf.data_ptr(), n_nodes, root, a.measure_map())
        b.sample(50000, seed=3)
        assert b.counts() == a.counts()


def test_packed_and_fused_histograms_hold_the_same_counts():
    """Registers of at most 32 classical bits are counted in packed 8-byte slots (key << 32 | count, one atomic per new
    outcome), wider ones in fused 16-byte slots; the option ``histogram_fused`` forces the second format."""
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import Engine
    for circ, shots in ((C.qft(20), 3000000), (C.random_circuit(12, 60, 2), 500000), (C.ghz(40), 100000)):
        with Engine(device=0, node_capacity=1 << 16) as e:
            e.load(circ)
            e.run(shots, 11)
            a = e.counts()
            e.set_option("histogram_fused", 1)
            e.run(shots, 11)
            assert e.counts() == a and sum(a.values()) == shots
            e.set_option("histogram_fused", 0)
            # accumulate in two halves: the same histogram
            e.apply()
            e.sample(shots // 2, seed=11, shot_offset=0, accumulate=False)
            e.sample(shots - shots // 2, seed=11, shot_offset=shots // 2, accumulate=True)
            assert e.counts() == a

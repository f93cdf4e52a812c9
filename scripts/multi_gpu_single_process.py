"""ONE process, several GPUs (run it with `gpurun --gpus N -- python scripts/multi_gpu_single_process.py`):
jkq_sample_multi_gpu and QasmSimulator(devices=...) against one GPU -- identical counts -- and the wall time of the
BASELINE configs[4] sampling sweep at 1 .. N devices.  Prints one JSON line."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from qiskit_jku_provider_b200 import circuits as C  # noqa: E402
from qiskit_jku_provider_b200.engine import Engine  # noqa: E402

n_gpu = torch.cuda.device_count()
out = {"gpus": n_gpu}
# (a) identical counts
for circ, shots in ((C.ghz(24), 1 << 22), (C.random_circuit(10, 60, 3), 2000001), (C.qft(14), 1 << 20)):
    with Engine(device=0, node_capacity=1 << 18) as e:
        e.load(circ)
        e.run(shots, 5)
        want = e.counts()
        e.apply()
        for g in range(1, n_gpu + 1):
            st = e.sample_multi(list(range(g)), shots, 5)
            assert e.counts() == want, (circ.name, g)
out["identical_counts"] = True
# (b) the sweep: wall = broadcast + sample + merge, through the C ABI (apply excluded: it runs on one GPU either way)
sweep = {}
for name, circ, shot_list in (("ghz24", C.ghz(24), (10 ** 8, 10 ** 9)), ("qft32", C.qft(32), (10 ** 8, 10 ** 9))):
    with Engine(device=0, node_capacity=1 << 18) as e:
        e.load(circ)
        e.apply()
        for shots in shot_list:
            base = None
            for g in sorted({1, 2, 4, 8, n_gpu} & set(range(1, n_gpu + 1))):
                e.sample_multi(list(range(g)), min(shots, 10 ** 7), 7)  # warm-up: helper engines, histograms, peer access
                best = None
                for _ in range(3):
                    st = e.sample_multi(list(range(g)), shots, 7)
                    if best is None or st["wall_ms"] < best["wall_ms"]:
                        best = st
                rec = {"shots_per_s_wall": shots / (best["wall_ms"] * 1e-3), "shots_per_s_kernel": shots / (best["sample_kernel_ms_max"] * 1e-3)}
                rec.update({k: best[k] for k in ("wall_ms", "broadcast_ms", "sample_wall_ms", "sample_kernel_ms_max", "merge_ms", "merged", "distinct_total")})
                if g == 1:
                    base = rec["shots_per_s_wall"]
                rec["wall_efficiency_vs_1"] = rec["shots_per_s_wall"] / (g * base)
                sweep["%s_%.0e_x%d" % (name, shots, g)] = rec
out["sweep"] = sweep
# (c) the experiments of one Qobj over the devices
from qiskit_jku_provider_b200 import QasmSimulator  # noqa: E402
from qiskit_jku_provider_b200.compat import ClassicalRegister, QuantumCircuit, QuantumRegister, assemble  # noqa: E402
import random  # noqa: E402
circs = []
for s in range(16):
    rnd = random.Random(s)
    q, c = QuantumRegister(12, "q"), ClassicalRegister(12, "c")
    qc = QuantumCircuit(q, c, name="c%d" % s)
    for _ in range(60):
        a, b = rnd.sample(range(12), 2)
        k = rnd.randrange(4)
        if k == 0:
            qc.h(q[a])
        elif k == 1:
            qc.cx(q[a], q[b])
        elif k == 2:
            qc.t(q[a])
        else:
            qc.u3(rnd.random(), rnd.random(), rnd.random(), q[a])
    qc.measure(q, c)
    circs.append(qc)
qobj = assemble(circs, shots=2048, seed=3)
one = QasmSimulator(silent=True, device=0)
r1 = one.run(qobj).result()
t0 = time.perf_counter(); r1 = one.run(qobj).result(); t_one = time.perf_counter() - t0
many = QasmSimulator(silent=True, devices=list(range(n_gpu)))
rn = many.run(qobj).result()
t0 = time.perf_counter(); rn = many.run(qobj).result(); t_many = time.perf_counter() - t0
for a, b in zip(r1.results, rn.results):
    assert a.data["counts"] == b.data["counts"]
out["qobj_16_experiments"] = {"one_gpu_s": t_one, "all_gpus_s": t_many}
print(json.dumps(out))

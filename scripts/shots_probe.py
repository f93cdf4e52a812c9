"""Per-shot re-simulation throughput: k_batch_shots (one launch) vs the sequential host loop, teleportation circuit."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from qiskit_jku_provider_b200.batch import BatchRunner
from qiskit_jku_provider_b200.engine import Engine
from qiskit_jku_provider_b200.qasm import parse_qasm
from test_qasm_frontend import load_case

prog = parse_qasm(load_case("teleport")[0])
t0 = time.perf_counter()
br = BatchRunner(workers=148, device=0)
t1 = time.perf_counter()
print("BatchRunner(148 engines) create: %.3f s" % (t1 - t0))
for shots in (1000, 10000, 100000, 1000000):
    br.run_shots(prog.n_qubits, prog.ops, min(shots, 1000), seed=1)
    t0 = time.perf_counter()
    keys, bits, status, gates, ms = br.run_shots(prog.n_qubits, prog.ops, shots, seed=1)
    w = time.perf_counter() - t0
    print("shots %8d  kernel %.2f ms  wall %.3f s  -> %.0f shots/s kernel, %.0f shots/s wall, %.1f gate applies/shot, failed %d" % (
        shots, ms, w, shots / (ms * 1e-3), shots / w, gates / shots, int(status.astype(bool).sum())))
eng = Engine(device=0)
eng.load_ops(prog.n_qubits, prog.ops)
t0 = time.perf_counter()
for s in range(300):
    eng.set_option("measure_seed", 100 + s)
    eng.apply()
    eng.sample(1, seed=1, shot_offset=s)
    eng.counts_arrays()
w = time.perf_counter() - t0
print("sequential host loop: %.0f shots/s" % (300 / w))

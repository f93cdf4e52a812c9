"""Host clock around the steps of one QasmSimulator.run(qobj).result().get_counts() call (QFT-32, 1024 shots)."""
import math, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from qiskit_jku_provider_b200 import QasmSimulator
from qiskit_jku_provider_b200 import qasm_simulator_jku as Q
from qiskit_jku_provider_b200.compat import ClassicalRegister, QuantumCircuit, QuantumRegister, assemble
from qiskit_jku_provider_b200.engine import Engine

qr, cr = QuantumRegister(32, "q"), ClassicalRegister(32, "c")
qc = QuantumCircuit(qr, cr, name="qft32")
for i in range(32):
    if i % 4 in (0, 2):
        qc.x(qr[i])
for i in range(32):
    qc.h(qr[i])
    for j in range(i + 1, 32):
        qc.cu1(math.pi / 2 ** (j - i), qr[j], qr[i])
qc.measure(qr, cr)
backend = QasmSimulator(silent=True, device=0, node_capacity=1 << 20)
qobj = assemble(qc, shots=1024, seed=1)
acc = {}
def wrap(obj, name):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            acc[name] = acc.get(name, 0.0) + time.perf_counter() - t0
    setattr(obj, name, g)
W = Q.JKUSimulatorWrapper
for n in ("compute_measurement_data", "convert_qobj_circuit_to_ops", "collect_output", "run_experiment"):
    wrap(W, n)
for n in ("load_ops", "run", "stats", "counts_arrays", "set_option"):
    wrap(Engine, n)
for _ in range(3):
    backend.run(qobj).result().get_counts()
acc.clear()
N = 20
t0 = time.perf_counter()
for _ in range(N):
    r = backend.run(qobj).result()
    t1 = time.perf_counter()
    r.get_counts()
    acc["get_counts"] = acc.get("get_counts", 0.0) + time.perf_counter() - t1
tot = time.perf_counter() - t0
print("ms per call %.3f" % (1e3 * tot / N))
for k, v in sorted(acc.items(), key=lambda kv: -kv[1]):
    print("  %-32s %.3f ms" % (k, 1e3 * v / N))
print("  kernel apply_ms %.3f sample_ms %.3f" % (backend._wrapper.engine.stats()["apply_ms"] if hasattr(backend, "_wrapper") else -1, -1))

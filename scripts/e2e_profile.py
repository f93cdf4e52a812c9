"""Where the host time of one QasmSimulator.run(qobj).result() call goes (QFT-32, 1024 shots): cProfile + phase timers."""
import cProfile, math, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from qiskit_jku_provider_b200 import QasmSimulator
from qiskit_jku_provider_b200.compat import ClassicalRegister, QuantumCircuit, QuantumRegister, assemble

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
for _ in range(3):
    backend.run(qobj).result().get_counts()
t0 = time.perf_counter()
for _ in range(10):
    r = backend.run(qobj).result()
    r.get_counts()
print("ms per call: %.3f" % (1e2 * (time.perf_counter() - t0)))
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    backend.run(qobj).result().get_counts()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

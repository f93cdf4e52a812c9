"""usage: batch_sweep.py [n_circuits [first_seed]]
A batch of random circuits (4-16 qubits, the deep few-qubit ones outgrow one SM's value table and fall back) through
k_batch against the one-at-a-time engine: identical counts, circuit by circuit."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.batch import BatchRunner
from qiskit_jku_provider_b200.engine import Engine

n_circ = int(sys.argv[1]) if len(sys.argv) > 1 else 300
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
circs = [C.random_circuit(4 + s % 13, 40 + 10 * (s % 6), s) for s in range(seed0, seed0 + n_circ)]
with BatchRunner(workers=148, big=dict(node_capacity=1 << 20)) as br:
    got, gates, ms = br.run(circs, shots=512, seed=7)
    again, _, _ = br.run(circs, shots=512, seed=7)
    fallbacks = br.fallbacks // 2
assert got == again, "two runs of the same batch differ"
assert gates == sum(len(c) for c in circs)
bad = []
with Engine(device=0, node_capacity=1 << 20) as eng:
    for i, c in enumerate(circs):
        eng.load(c)
        eng.run(512, 7 + i)
        if eng.counts() != got[i]:
            bad.append(c.name)
assert not bad, bad
print("batch sweep ok: %d circuits, %d gate applies, %d fell back to the full engine, %.1f ms on the device" % (n_circ, gates, fallbacks, ms))

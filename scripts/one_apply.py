"""One QFT apply (for ncu): python scripts/one_apply.py [n_qubits] [reps]"""
import sys

sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = Engine(device=0, node_capacity=1 << 20)
eng.load(C.qft(n))
for _ in range(reps):
    eng.apply()
print("apply_ms", eng.stats()["apply_ms"])
eng.close()

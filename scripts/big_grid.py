"""random n=22, 200 ops (43 M nodes created): the grid path at large frontiers.  usage: big_grid.py [n ops]"""
import sys
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine
n, ops = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (22, 200)
eng = Engine(device=0, node_capacity=1 << 25, value_capacity=1 << 26)
eng.load(C.random_circuit(n, ops, 1))
eng.apply()
st = eng.stats()
print("apply_ms", st["apply_ms"], "items", st["items"], "nodes_created", st["nodes_created"], "gc", st["gc_runs"], "retry", st["retry_rounds"],
      "Mitems/s", st["items"] / st["apply_ms"] / 1e3, "GB/s algorithmic", st["items"] * 320 / st["apply_ms"] / 1e6)

"""Force an engine error inside the grid kernel, then keep using the handle."""
import sys
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine
eng = Engine(device=0, node_capacity=1 << 22, value_capacity=1 << 14)
for circ in (C.random_circuit(18, 120, 1), C.random_circuit(18, 160, 1), C.ghz(10), C.random_circuit(18, 100, 1)):
    try:
        eng.load(circ)
        eng.apply()
        print(circ.name, "ok", eng.stats()["nodes_created"], flush=True)
    except Exception as e:
        print(circ.name, "ERR", str(e)[:200], flush=True)

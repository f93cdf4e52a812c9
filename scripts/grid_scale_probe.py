"""Grid-path scaling: blown-up random circuits of growing size."""
import json
import sys
import time
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine

cases = [(18, 120), (18, 140), (18, 160), (20, 160), (20, 200), (22, 200)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
import os
eng = Engine(device=0, node_capacity=1 << int(os.environ.get('NODE_LOG2', '25')), value_capacity=1 << int(os.environ.get('VAL_LOG2', '22')))
for n, ops in cases:
    circ = C.random_circuit(n, ops, 1)
    eng.load(circ)
    t = time.time()
    try:
        eng.apply()
    except Exception as e:
        print(n, ops, "ERR", e, flush=True)
        continue
    w = time.time() - t
    st = eng.stats()
    print(json.dumps({"name": circ.name, "applies": len(circ), "apply_ms": round(st["apply_ms"], 2), "wall_s": round(w, 3), "items": st["items"],
                      "nodes_created": st["nodes_created"], "live_nodes": st["live_nodes"], "max_frontier": st["max_frontier"],
                      "Mitems_per_s": round(st["items"] / st["apply_ms"] / 1e3, 2), "gc": st["gc_runs"], "retry": st["retry_rounds"],
                      "vals": st["live_values"], "grid": st["grid_launches"]}), flush=True)

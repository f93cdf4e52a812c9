"""solo_max on circuits that reach the grid path.  usage: solo_probe2.py [solo_max,...]"""
import json, sys
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine

vals = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 16, 8, 4, 2]
circs = [C.random_circuit(18, 100, 1), C.random_circuit(18, 120, 1), C.random_circuit(18, 160, 1), C.random_circuit(18, 40, 170), C.random_circuit(18, 40, 738)]
eng = Engine(device=0, node_capacity=1 << 20)
for sm in vals:
    eng.set_option("solo_max", sm)
    for c in circs:
        eng.load(c)
        best = None
        for _ in range(2):
            eng.apply()
            st = eng.stats()
            if best is None or st["apply_ms"] < best["apply_ms"]:
                best = st
        print(json.dumps({"solo_max": sm, "name": c.name, "apply_ms": round(best["apply_ms"], 3), "us_per_gate": round(1e3 * best["apply_ms"] / len(c), 2),
                          "retry": best["retry_rounds"], "cta": best["cta_launches"], "grid": best["grid_launches"]}), flush=True)
eng.close()

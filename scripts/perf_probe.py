"""Quick perf probe: apply time per gate for the BASELINE shapes.
usage: perf_probe.py [threads,threads,...] [name-substring]"""
import json
import sys

sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine

threads = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [256]
only = sys.argv[2] if len(sys.argv) > 2 else ""
circs = [C.qft(32), C.ghz(24), C.grover(20, 20), C.random_circuit(18, 40, 1), C.random_circuit(18, 100, 1), C.random_circuit(18, 120, 1)]
for th in threads:
    eng = Engine(device=0, node_capacity=1 << 20, cta_threads=th)
    for c in circs:
        if only and only not in c.name:
            continue
        eng.load(c)
        best = None
        for _ in range(3):
            eng.apply()
            st = eng.stats()
            if best is None or st["apply_ms"] < best["apply_ms"]:
                best = st
        print(json.dumps({"threads": th, "name": c.name, "applies": len(c), "apply_ms": round(best["apply_ms"], 3),
                          "us_per_gate": round(1e3 * best["apply_ms"] / len(c), 2), "items": best["items"], "thin": best["thin_gates"],
                          "retry": best["retry_rounds"], "nodes_created": best["nodes_created"], "gc": best["gc_runs"],
                          "cta": best["cta_launches"], "grid": best["grid_launches"]}), flush=True)
    eng.close()

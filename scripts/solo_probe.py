"""How wide a level should the cursor warp run by itself?  usage: solo_probe.py [solo_max,...]  (0 = one warp's width)"""
import json, sys
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine

vals = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [0, 8, 2, 1]
circs = [C.qft(32), C.grover(20, 20), C.random_circuit(18, 40, 1), C.random_circuit(18, 100, 1)]
eng = Engine(device=0, node_capacity=1 << 20)
for wide in (0, 1):
    for sm in vals:
        if not wide and sm:
            continue
        eng.set_option("cta_always_wide", wide)
        eng.set_option("solo_max", sm)
        for c in circs:
            eng.load(c)
            best = None
            for _ in range(3):
                eng.apply()
                st = eng.stats()
                if best is None or st["apply_ms"] < best["apply_ms"]:
                    best = st
            print(json.dumps({"always_wide": wide, "solo_max": sm, "name": c.name, "apply_ms": round(best["apply_ms"], 3),
                              "us_per_gate": round(1e3 * best["apply_ms"] / len(c), 2), "retry": best["retry_rounds"], "cta": best["cta_launches"]}), flush=True)
eng.close()

"""A/B of library builds on circuits that reach the grid path.  usage: grid_ab.py lib1.so[,lib2.so...] [big]
Each build runs in its own process (JKQ_LIB picks the library); best of three applies per circuit."""
import json, os, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CHILD = r'''
import json, sys
sys.path.insert(0, %r)
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine
big = %r
circs = [C.random_circuit(18, 100, 1), C.random_circuit(18, 120, 1), C.random_circuit(18, 160, 1), C.random_circuit(18, 40, 170), C.random_circuit(18, 40, 738)]
if big: circs.append(C.random_circuit(22, 200, 1))
eng = Engine(device=0, node_capacity=1 << (25 if big else 20), value_capacity=1 << (26 if big else 20))
for c in circs:
    eng.load(c)
    best = None
    for _ in range(3):
        eng.apply()
        st = eng.stats()
        if best is None or st["apply_ms"] < best["apply_ms"]: best = st
    print(json.dumps({"name": c.name, "apply_ms": round(best["apply_ms"], 3), "items": best["items"], "Mitems_s": round(best["items"] / best["apply_ms"] / 1e3, 2),
                      "norm_err": abs(best["norm"] - 1.0), "gc": best["gc_runs"], "retry": best["retry_rounds"]}), flush=True)
eng.close()
'''
libs = sys.argv[1].split(",")
big = len(sys.argv) > 2
for lib in libs:
    print("==", lib, flush=True)
    env = dict(os.environ, JKQ_LIB=os.path.abspath(lib))
    subprocess.run([sys.executable, "-c", CHILD % (os.path.abspath(ROOT), big)], env=env, check=False)

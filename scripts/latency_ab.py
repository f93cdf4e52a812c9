"""A/B of library builds on the latency-bound shapes (default launch shapes) and the two sampling diagrams.
usage: latency_ab.py lib1.so[,lib2.so...]   Each build runs in its own process (JKQ_LIB picks the library)."""
import os, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
CHILD = r'''
import json, sys, time
sys.path.insert(0, %r)
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine
eng = Engine(device=0, node_capacity=1 << 20)
for c in (C.qft(32), C.grover(20, 20), C.ghz(24), C.random_circuit(18, 40, 1)):
    eng.load(c)
    best = None
    for _ in range(7):
        eng.apply()
        st = eng.stats()
        if best is None or st["apply_ms"] < best["apply_ms"]: best = st
    print(json.dumps({"name": c.name, "apply_ms": round(best["apply_ms"], 3), "us_per_gate": round(1e3 * best["apply_ms"] / len(c), 3), "thin": best["thin_gates"]}), flush=True)
for c, shots in ((C.ghz(24), 10 ** 9), (C.qft(32), 10 ** 8)):
    eng.load(c); eng.apply()
    best = None
    for _ in range(3):
        eng.sample(shots, seed=7)
        ms = eng.stats()["sample_kernel_ms"]
        best = ms if best is None or ms < best else best
    print(json.dumps({"sample": c.name, "shots": shots, "ms": round(best, 3), "Gshots_s": round(shots / best / 1e6, 2), "distinct": eng.distinct_outcomes()}), flush=True)
eng.close()
'''
for lib in sys.argv[1].split(","):
    print("==", lib, flush=True)
    subprocess.run([sys.executable, "-c", CHILD % os.path.abspath(ROOT)], env=dict(os.environ, JKQ_LIB=os.path.abspath(lib)), check=False)

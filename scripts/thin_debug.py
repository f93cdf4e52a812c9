"""Bisect thin-gate failures: random circuits against a dense simulation under each thin_off mask.
usage: thin_debug.py seed0 nseeds"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine
from parity_sweep_lib import dense

seed0 = int(sys.argv[1]); ns = int(sys.argv[2])
with Engine(device=0, node_capacity=1 << 20) as eng:
    for seed in range(seed0, seed0 + ns):
        n = 6 + seed % 7
        circ = C.random_circuit(n, 40 + 10 * (seed % 6), seed)
        want = dense(circ)
        for off in (8, 7, 6, 5, 3, 4, 2, 1, 0):
            eng.set_option("thin_off", off)
            eng.set_option("force_mode", 0)
            eng.load(circ)
            try:
                eng.apply()
                err = float(np.abs(eng.statevector() - want).max())
                st = eng.stats()
                print("seed", seed, "n", n, "gates", len(circ), "off", off, "err %.2e" % err, "thin", st["thin_gates"], "FAIL" if err > 1e-9 else "", flush=True)
            except Exception as e:
                print("seed", seed, "n", n, "off", off, "EXC", str(e)[:100], flush=True)

"""First gate at which the thin path deviates from a dense simulation.  usage: thin_debug2.py seed"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine, circuit_ops
from parity_sweep_lib import dense

seed = int(sys.argv[1])
n = 6 + seed % 7
circ = C.random_circuit(n, 40 + 10 * (seed % 6), seed)
ops = circuit_ops(circ)
print("n", n, "ops", len(ops))
with Engine(device=0, node_capacity=1 << 20) as eng:
    lo, hi = 0, len(ops)
    def bad(k):
        sub = C.Circuit(n, circ.ops[:k], "sub")
        eng.set_option("thin_off", 0)
        eng.load(sub)
        try:
            eng.apply()
            err = float(np.abs(eng.statevector() - dense(sub)).max())
        except Exception as e:
            print(k, "EXC", str(e)[:80]); return True
        return err > 1e-9
    while hi - lo > 1:
        mid = (lo + hi) // 2
        if bad(mid): hi = mid
        else: lo = mid
    print("first bad prefix length", hi)
    g = ops[hi - 1]
    print("gate", hi - 1, "target", g["target"], "cpos", bin(int(g["ctrl_pos"])), "cneg", bin(int(g["ctrl_neg"])), "m", np.round(np.array(g["m"]), 4))
    for off in (8, 0):
        sub = C.Circuit(n, circ.ops[:hi], "sub")
        eng.set_option("thin_off", off); eng.set_option("level_log", 1)
        eng.load(sub)
        try:
            eng.apply()
            print("off", off, "levels", eng.level_log(hi)[-2:], "thin", eng.stats()["thin_gates"])
        except Exception as e:
            print("off", off, "EXC", e)

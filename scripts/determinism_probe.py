"""Run one blown-up circuit several times: level-count sums and amplitude error against the oracle."""
import sys
import numpy as np
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine
import oracle

circ = C.random_circuit(18, int(sys.argv[1]) if len(sys.argv) > 1 else 100, 1)
o = oracle.run_oracle(circ, amps=True, extra=["--amp-limit", "65536", "--tie-tol"])
oa = np.array([complex(*x) for x in o["amps"]])
print("oracle", sum(o["level_counts"]), o["level_counts"])
for mode, th in ((0, 64), (0, 256), (2, 64), (1, 32)):
    for rep in range(3):
        eng = Engine(device=0, node_capacity=1 << 20, cta_threads=th)
        eng.set_option("force_mode", mode)
        eng.load(circ)
        try:
            eng.apply()
            lc = eng.level_counts()
            err = np.abs(eng.amplitudes(0, 65536) - oa).max()
            st = eng.stats()
            print("mode", mode, "threads", th, "sum", sum(lc), lc[:6], "amp_err %.2e" % err, "norm-1 %.1e" % (st["norm"] - 1), "vals", st["live_values"], "grid", st["grid_launches"])
        except Exception as e:
            print("mode", mode, "threads", th, "ERR", e)
        eng.close()

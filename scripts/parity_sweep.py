"""usage: parity_sweep.py [n_seeds [n_min [n_span [first_seed]]]]
Wider parity sweep than the test suite: random circuits through the single-CTA path, the grid path and the automatic
mode against a dense NumPy simulation of the same ops (amplitudes to 1e-9), plus run-to-run determinism of the counts."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine, circuit_ops


def dense(circ):
    n = circ.n
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1.0
    idx = np.arange(1 << n)
    for op in circuit_ops(circ):
        t = int(op["target"])
        m = np.array(op["m"]).view(np.complex128).reshape(2, 2)
        cp, cn = int(op["ctrl_pos"]), int(op["ctrl_neg"])
        lo = idx[((idx & cp) == cp) & ((idx & cn) == 0) & (((idx >> t) & 1) == 0)]
        hi = lo | (1 << t)
        a, b = psi[lo].copy(), psi[hi].copy()
        psi[lo] = m[0, 0] * a + m[0, 1] * b
        psi[hi] = m[1, 0] * a + m[1, 1] * b
    return psi


worst = 0.0
n_checked = 0
with Engine(device=0, node_capacity=1 << 20) as eng:
    for kv in filter(None, os.environ.get("SWEEP_OPTS", "").split(",")):  # e.g. SWEEP_OPTS=cta_val_limit=300,cta_node_limit=200
        k, v = kv.split("=")
        eng.set_option(k, int(v))
    n_min = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    n_span = int(sys.argv[3]) if len(sys.argv) > 3 else 7
    seed0 = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    for seed in range(seed0, seed0 + (int(sys.argv[1]) if len(sys.argv) > 1 else 40)):
        n = n_min + seed % n_span
        circ = C.random_circuit(n, 40 + 10 * (seed % 6), seed)
        want = dense(circ)
        for mode in (0, 1, 2):
            if os.environ.get("SWEEP_VERBOSE"):
                print("seed", seed, "n", n, "gates", len(circ), "mode", mode, flush=True)
            eng.set_option("force_mode", mode)
            eng.load(circ)
            try:
                eng.run(2000, seed)
            except Exception as e:  # force_mode=1 on a circuit that outgrows one SM
                assert mode == 1, (seed, mode, e)
                continue
            got = eng.statevector()
            err = float(np.abs(got - want).max())
            worst = max(worst, err)
            assert err < 1e-9, (seed, mode, err)
            k1 = eng.counts()
            eng.run(2000, seed)
            assert eng.counts() == k1, (seed, mode, "counts differ between two runs")
            n_checked += 1
    eng.set_option("force_mode", 0)
    for circ in (C.qft(16), C.qft(24), C.grover(10, 12, marked=0x2A5), C.ghz(30)):
        eng.load(circ)
        eng.apply()
        if circ.n <= 16:
            err = float(np.abs(eng.statevector() - dense(circ)).max())
            worst = max(worst, err)
            assert err < 1e-9, (circ.name, err)
        assert abs(eng.stats()["norm"] - 1.0) < 1e-9, circ.name
print("parity sweep ok: %d runs, worst amplitude error %.2e" % (n_checked, worst))

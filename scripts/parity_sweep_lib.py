"""dense NumPy simulation of an op list (shared by the sweep scripts)"""
import numpy as np
from qiskit_jku_provider_b200.engine import circuit_ops


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

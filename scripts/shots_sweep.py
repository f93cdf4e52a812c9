"""usage: shots_sweep.py [n_programs [first_seed]]
Random OpenQASM programs with mid-circuit measure / reset / if: the per-shot kernel (jkq_batch_run_shots, one CTA per shot)
against the sequential host-driven path (jkq_apply + jkq_sample per shot) -- same Philox streams, so every shot must give
the same final outcome and the same mid-circuit bits; projective and physical reset."""
import os, random, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_gpu_shots import sequential
from test_qasm_frontend import HEADER
from qiskit_jku_provider_b200.batch import BatchRunner
from qiskit_jku_provider_b200.engine import Engine
from qiskit_jku_provider_b200.qasm import parse_qasm
from qiskit_jku_provider_b200.jkusimulatorerror import JKUSimulatorError


def program(seed):
    r = random.Random(seed)
    n = r.randint(3, 8)
    lines = ["qreg q[%d]; creg c[2]; creg d[%d];" % (n, n)]
    for _ in range(r.randint(15, 60)):
        k = r.random()
        a, b, t = r.sample(range(n), 3)
        g1 = r.choice(["h q[%d];", "x q[%d];", "t q[%d];", "ry(%.3f) q[%%d];" % r.uniform(0.1, 3.0), "rz(%.3f) q[%%d];" % r.uniform(0.1, 3.0),
                       "u3(%.3f,%.3f,%.3f) q[%%d];" % (r.uniform(0, 3), r.uniform(0, 3), r.uniform(0, 3))])
        if k < 0.35: s = g1 % a
        elif k < 0.60: s = "cx q[%d], q[%d];" % (a, b)
        elif k < 0.68: s = "ccx q[%d], q[%d], q[%d];" % (a, b, t)
        elif k < 0.80: s = "measure q[%d] -> c[%d];" % (a, r.randint(0, 1))
        elif k < 0.85: s = "measure q[%d] -> d[%d];" % (a, a)
        elif k < 0.90: s = "reset q[%d];" % a
        else: s = "if (c == %d) %s" % (r.randint(0, 3), r.choice([g1 % a, "cx q[%d], q[%d];" % (a, b), "measure q[%d] -> d[%d];" % (a, a), "reset q[%d];" % a]))
        lines.append(s)
    return HEADER + "\n".join(lines) + "\n"


import numpy as np
from qiskit_jku_provider_b200.engine import OP_GATE, OP_MEASURE_ONE, OP_RESET, OP_IF


def trajectories(ops, n, physical):
    """Every (final state, classical bits) a dense simulation of the program can end in: branches at each mid-circuit
    measurement (and at each physical reset).  Independent of the engine: plain NumPy on 2^n amplitudes."""
    idx = np.arange(1 << n)
    out = []

    def project(psi, t, o):
        keep = ((idx >> t) & 1) == o
        p = float((np.abs(psi[keep]) ** 2).sum())
        if p < 1e-13: return None, p
        q = np.where(keep, psi, 0) / np.sqrt(p)
        return q, p

    def flip(psi, t):
        return psi[idx ^ (1 << t)]

    def run(i, psi, cb, skip_until):
        while i < len(ops):
            op = ops[i]; kind = int(op["kind"]); t = int(op["target"])
            if i < skip_until: i += 1; continue
            if kind == OP_IF:
                span = int(op["arg"]) or 1
                if (cb & int(op["ctrl_pos"])) != int(op["ctrl_neg"]): skip_until = i + 1 + span
                i += 1; continue
            if kind == OP_GATE:
                m = np.array(op["m"]).view(np.complex128).reshape(2, 2); cp, cn = int(op["ctrl_pos"]), int(op["ctrl_neg"])
                lo = idx[((idx & cp) == cp) & ((idx & cn) == 0) & (((idx >> t) & 1) == 0)]; hi = lo | (1 << t)
                a, b = psi[lo].copy(), psi[hi].copy(); psi = psi.copy(); psi[lo] = m[0, 0] * a + m[0, 1] * b; psi[hi] = m[1, 0] * a + m[1, 1] * b
            elif kind == OP_MEASURE_ONE:
                for o in (0, 1):
                    q, p = project(psi, t, o)
                    if q is not None: run(i + 1, q, (cb & ~(1 << int(op["arg"]))) | (o << int(op["arg"])), skip_until)
                return
            elif kind == OP_RESET:
                if physical:
                    for o in (0, 1):
                        q, p = project(psi, t, o)
                        if q is not None: run(i + 1, flip(q, t) if o else q, cb, skip_until)
                    return
                q, p = project(psi, t, 0)
                if q is None: q, p = project(flip(psi, t), t, 0)
                psi = q
            i += 1
        out.append((psi, cb))

    psi0 = np.zeros(1 << n, dtype=np.complex128); psi0[0] = 1
    run(0, psi0, 0, 0)
    return out


n_prog = int(sys.argv[1]) if len(sys.argv) > 1 else 40
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
shots = 48
checked = skipped = 0
worst = 0.0
with BatchRunner(workers=148, device=0) as br, Engine(device=0, node_capacity=1 << 18) as eng:
    for seed in range(seed0, seed0 + n_prog):
        text = program(seed)
        prog = parse_qasm(text)
        for physical in (False, True):
            try:
                want_keys, want_bits = sequential(eng, prog, shots, 5 + seed, 99991 * (seed + 1), physical)
            except JKUSimulatorError as e:  # e.g. a projective reset of a qubit that is surely 1: "0-vector", like the reference
                skipped += 1
                continue
            keys, bits, status, gates, ms = br.run_shots(prog.n_qubits, prog.ops, shots, seed=5 + seed, measure_seed0=99991 * (seed + 1), physical_reset=physical)
            assert not status.any(), (seed, physical, "status")
            assert [int(k) for k in keys] == want_keys, (seed, physical, "keys", text)
            assert [int(b) for b in bits] == want_bits, (seed, physical, "bits", text)
            # the engine still holds the last shot of the sequential path: its state must be one of the dense trajectories
            sv, cb = eng.statevector(), want_bits[-1]
            traj = trajectories(prog.ops, prog.n_qubits, physical)
            best = min((float(np.abs(sv - psi).max()) for psi, b in traj if b == cb), default=9.0)
            assert best < 1e-9, (seed, physical, "no dense trajectory ends in the engine's state", best, len(traj), text)
            worst = max(worst, best)
            checked += 1
print("shots sweep ok: %d program x reset-mode runs of %d shots each (%d skipped: the sequential path reported an error); final state vs dense trajectory: worst %.2e" % (checked, shots, skipped, worst))

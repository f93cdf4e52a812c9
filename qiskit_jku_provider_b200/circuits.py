"""Circuit description shared by the CUDA engine, the oracle and the reference harness.

A circuit is a list of *primitive* operations -- exactly the two primitives the reference
engine has, ``U(theta,phi,lambda)`` on one target and ``CX`` (reference
``src/QASMsimulator.cpp:281-357``), plus the multi-controlled 2x2 gate the engine-level builder
``QMDDmvlgate`` accepts (``src/QMDDpackage.cpp:1746-1828``).  Every named gate the backend
advertises is macro-expanded to those primitives the way the reference's ``qelib1`` string does
(``qiskit_jku_provider/qasm_simulator_jku.py:32-60``); one expanded primitive == one ``ApplyGate``
(``src/Simulator.cpp:466``) == one "gate apply" of the benchmark metric.

Angles keep a symbolic multiple of pi when they have one, so the reference harness can evaluate
``pi/2**k`` in 200-bit arithmetic like the reference's OpenQASM parser does
(``src/QASMsimulator.cpp:127``), while the CUDA engine and the FP64 oracle get the nearest double.

This module also holds the workload generators of BASELINE.json's configs (GHZ, QFT, Grover,
random circuits with the semantics of the reference's ``test/_random_circuit_generator.py``).
"""

import math
import random
from fractions import Fraction


class Angle:
    """``value = const + pi * frac`` -- at most one of the two parts is non-zero in practice."""

    __slots__ = ("const", "frac")

    def __init__(self, const=0.0, frac=Fraction(0)):
        self.const = float(const)
        self.frac = Fraction(frac)

    @staticmethod
    def of(x):
        if isinstance(x, Angle):
            return x
        return Angle(float(x))

    @staticmethod
    def pi(num=1, den=1):
        return Angle(0.0, Fraction(num, den))

    def __float__(self):
        if self.frac == 0:
            return self.const
        v = math.pi * self.frac.numerator / self.frac.denominator
        return v + self.const if self.const else v

    def __neg__(self):
        return Angle(-self.const, -self.frac)

    def __add__(self, o):
        o = Angle.of(o)
        return Angle(self.const + o.const, self.frac + o.frac)

    __radd__ = __add__

    def __sub__(self, o):
        return self + (-Angle.of(o))

    def __rsub__(self, o):
        return Angle.of(o) - self

    def __truediv__(self, k):
        return Angle(self.const / k, self.frac / k)

    def __mul__(self, k):
        return Angle(self.const * k, self.frac * k)

    __rmul__ = __mul__

    def token(self):
        """Text form understood by oracle/ref_harness.cpp and oracle/qmdd_oracle.c."""
        if self.frac == 0 or self.const != 0.0:
            return repr(float(self))
        num, den = self.frac.numerator, self.frac.denominator
        s = "-pi" if num < 0 else "pi"
        if abs(num) != 1:
            s += "*%d" % abs(num)
        if den != 1:
            s += "/%d" % den
        return s

    def __repr__(self):
        return "Angle(%s)" % self.token()


ZERO = Angle(0.0)
PI = Angle.pi()


class Op:
    """One primitive.  kind: 'u' (theta,phi,lam, optional controls) or 'g' (named exact matrix:
    'x' = the reference's Nm, 'z' = Zm, 'h' = Hm, 's' = Sm; src/QMDDpackage.cpp:37-72)."""

    __slots__ = ("kind", "target", "theta", "phi", "lam", "name", "controls")

    def __init__(self, kind, target, theta=ZERO, phi=ZERO, lam=ZERO, name="", controls=()):
        self.kind = kind
        self.target = int(target)
        self.theta, self.phi, self.lam = Angle.of(theta), Angle.of(phi), Angle.of(lam)
        self.name = name
        self.controls = tuple((int(q), int(v)) for q, v in controls)

    def line(self):
        ctl = " ".join(("%d" % q) if v else ("~%d" % q) for q, v in self.controls)
        if self.kind == "u":
            ang = "%s %s %s" % (self.theta.token(), self.phi.token(), self.lam.token())
            if not self.controls:
                return "u %d %s" % (self.target, ang)
            return "mcu %d %s %d %s" % (self.target, ang, len(self.controls), ctl)
        if self.name == "x" and len(self.controls) == 1 and self.controls[0][1] == 1:
            return "cx %d %d" % (self.controls[0][0], self.target)
        return ("mcg %s %d %d %s" % (self.name, self.target, len(self.controls), ctl)).rstrip()

    def qasm(self):
        """OpenQASM 2.0 statement (builtin U / CX only); ValueError for a primitive QASM cannot express."""
        if self.kind == "u" and not self.controls:
            return "U(%s,%s,%s) q[%d];" % (self.theta.token(), self.phi.token(), self.lam.token(), self.target)
        if self.kind == "g" and self.name == "x" and len(self.controls) == 1 and self.controls[0][1] == 1:
            return "CX q[%d],q[%d];" % (self.controls[0][0], self.target)
        raise ValueError("no OpenQASM 2.0 form for %s" % self.line())

    def matrix(self):
        """2x2 complex matrix in FP64.  U follows the SU(2) phase convention of
        src/QASMsimulator.cpp:295-298 (x = U(pi,0,pi) = -iX, u1(l) = diag(e^{-il/2}, e^{il/2}))."""
        if self.kind == "g":
            h = (0 + 1 * math.sqrt(2.0)) / 2  # Qmake(0,1,2), src/QMDDcomplexD.cpp:297-306
            return {
                "x": ((0j, 1 + 0j), (1 + 0j, 0j)),
                "z": ((1 + 0j, 0j), (0j, -1 + 0j)),
                "h": ((h + 0j, h + 0j), (h + 0j, -h + 0j)),
                "s": ((1 + 0j, 0j), (0j, 1j)),
            }[self.name]
        th, ph, la = float(self.theta), float(self.phi), float(self.lam)
        c, s = math.cos(th / 2), math.sin(th / 2)
        return (
            (complex(math.cos(-(ph + la) / 2) * c, math.sin(-(ph + la) / 2) * c),
             complex(-math.cos(-(ph - la) / 2) * s, -math.sin(-(ph - la) / 2) * s)),
            (complex(math.cos((ph - la) / 2) * s, math.sin((ph - la) / 2) * s),
             complex(math.cos((ph + la) / 2) * c, math.sin((ph + la) / 2) * c)),
        )


def U(t, theta, phi, lam):
    return Op("u", t, theta, phi, lam)


def CX(c, t):
    return Op("g", t, name="x", controls=((c, 1),))


def MCG(name, t, controls):
    return Op("g", t, name=name, controls=tuple((q, 1) if isinstance(q, int) else q for q in controls))


def MCU(t, theta, phi, lam, controls):
    return Op("u", t, theta, phi, lam, controls=tuple((q, 1) if isinstance(q, int) else q for q in controls))


# --------------------------------------------------------------------------- qelib1 expansion
# Restates the gate bodies of the reference's qelib1 (qasm_simulator_jku.py:32-60).

def _u3(q, th, ph, la):
    return [U(q, th, ph, la)]


def _u2(q, ph, la):
    return [U(q, PI / 2, ph, la)]


def _u1(q, la):
    return [U(q, ZERO, ZERO, la)]


def _h(q):
    return _u2(q, ZERO, PI)


def _x(q):
    return _u3(q, PI, ZERO, PI)


def _s(q):
    return _u1(q, PI / 2)


def _sdg(q):
    return _u1(q, -PI / 2)


def _t(q):
    return _u1(q, PI / 4)


def _tdg(q):
    return _u1(q, -PI / 4)


def _ccx(a, b, c):
    return (_h(c) + [CX(b, c)] + _tdg(c) + [CX(a, c)] + _t(c) + [CX(b, c)] + _tdg(c) + [CX(a, c)]
            + _t(b) + _t(c) + _h(c) + [CX(a, b)] + _t(a) + _tdg(b) + [CX(a, b)])


def expand(name, qubits, params=()):
    """Named gate -> list of primitive Ops.  Raises KeyError for an unknown gate."""
    p = [Angle.of(x) for x in params]
    q = list(qubits)
    if name in ("u3", "U"):
        return _u3(q[0], p[0], p[1], p[2])
    if name == "u2":
        return _u2(q[0], p[0], p[1])
    if name == "u1":
        return _u1(q[0], p[0])
    if name in ("cx", "CX"):
        return [CX(q[0], q[1])]
    if name in ("id", "iden", "u0"):
        return _u3(q[0], ZERO, ZERO, ZERO)
    if name == "x":
        return _x(q[0])
    if name == "y":
        return _u3(q[0], PI, PI / 2, PI / 2)
    if name == "z":
        return _u1(q[0], PI)
    if name == "h":
        return _h(q[0])
    if name == "s":
        return _s(q[0])
    if name == "sdg":
        return _sdg(q[0])
    if name == "t":
        return _t(q[0])
    if name == "tdg":
        return _tdg(q[0])
    if name == "rx":
        return _u3(q[0], p[0], -PI / 2, PI / 2)
    if name == "ry":
        return _u3(q[0], p[0], ZERO, ZERO)
    if name == "rz":
        return _u1(q[0], p[0])
    if name == "cz":
        return _h(q[1]) + [CX(q[0], q[1])] + _h(q[1])
    if name == "cy":
        return _sdg(q[1]) + [CX(q[0], q[1])] + _s(q[1])
    if name == "swap":
        return [CX(q[0], q[1]), CX(q[1], q[0]), CX(q[0], q[1])]
    if name == "ch":
        a, b = q
        return (_h(b) + _sdg(b) + [CX(a, b)] + _h(b) + _t(b) + [CX(a, b)] + _t(b) + _h(b) + _s(b)
                + _x(b) + _s(a))
    if name == "ccx":
        return _ccx(*q)
    if name == "cswap":
        a, b, c = q
        return [CX(c, b)] + _ccx(a, b, c) + [CX(c, b)]
    if name == "crz":
        a, b = q
        return _u1(b, p[0] / 2) + [CX(a, b)] + _u1(b, -p[0] / 2) + [CX(a, b)]
    if name == "cu1":
        a, b = q
        return _u1(a, p[0] / 2) + [CX(a, b)] + _u1(b, -p[0] / 2) + [CX(a, b)] + _u1(b, p[0] / 2)
    if name == "cu3":
        c, t = q
        th, ph, la = p
        return (_u1(t, (la - ph) / 2) + [CX(c, t)] + _u3(t, -th / 2, ZERO, -(ph + la) / 2)
                + [CX(c, t)] + _u3(t, th / 2, ph, ZERO))
    if name == "rzz":
        a, b = q
        return [CX(a, b)] + _u1(b, p[0]) + [CX(a, b)]
    raise KeyError(name)


class Circuit:
    """n qubits + a list of primitive ops."""

    def __init__(self, n, ops=None, name="circuit"):
        self.n = int(n)
        self.ops = list(ops or [])
        self.name = name

    def add(self, name, qubits, params=()):
        self.ops.extend(expand(name, qubits, params))
        return self

    def text(self):
        return "qubits %d\n%s\n" % (self.n, "\n".join(op.line() for op in self.ops))

    def write(self, path):
        with open(path, "w") as f:
            f.write(self.text())
        return path

    def qasm(self, measure=False):
        """The circuit as an OpenQASM 2.0 program of builtin U / CX statements (what the reference's front end
        reduces every program to, QASMsimulator.cpp:281-357)."""
        lines = ["OPENQASM 2.0;", "qreg q[%d];" % self.n]
        if measure:
            lines.append("creg c[%d];" % self.n)
        lines += [op.qasm() for op in self.ops]
        if measure:
            lines.append("measure q -> c;")
        return "\n".join(lines) + "\n"

    def __len__(self):
        return len(self.ops)


# --------------------------------------------------------------------------- workloads

def ghz(n, star=False):
    """BASELINE config 1: h q0; cx chain (or the star of the reference's test_entangle)."""
    c = Circuit(n, name="ghz%d" % n)
    c.add("h", [0])
    for i in range(n - 1):
        c.add("cx", [0, i + 1] if star else [i, i + 1])
    return c


def qft(n, basis_state="1010"):
    """BASELINE config 2: x on even qubits (|1010...>), then the textbook QFT with cu1 expanded
    through qelib1 (5 primitives each).  n=32 -> 16 + 32 + 496*5 = 2528 applies."""
    c = Circuit(n, name="qft%d" % n)
    for i in range(n):
        if basis_state[i % len(basis_state)] == "1":
            c.add("x", [i])
    for i in range(n):
        c.add("h", [i])
        for j in range(i + 1, n):
            c.add("cu1", [j, i], [Angle.pi(1, 2 ** (j - i))])
    return c


def grover(n, iterations, marked=0x5A5A5):
    """BASELINE config 3: no ancilla, native (n-1)-controlled Z on the last qubit
    (reference: QMDDmvlgate(Zm, n, line)); 102 applies per iteration at n = 20."""
    c = Circuit(n, name="grover%d_%d" % (n, iterations))
    marked &= (1 << n) - 1
    ctl = list(range(n - 1))
    for q in range(n):
        c.add("h", [q])
    for _ in range(iterations):
        zeros = [q for q in range(n) if not (marked >> q) & 1]
        for q in zeros:
            c.add("x", [q])
        c.ops.append(MCG("z", n - 1, ctl))
        for q in zeros:
            c.add("x", [q])
        for q in range(n):
            c.add("h", [q])
        for q in range(n):
            c.add("x", [q])
        c.ops.append(MCG("z", n - 1, ctl))
        for q in range(n):
            c.add("x", [q])
        for q in range(n):
            c.add("h", [q])
    return c


# op_signature of the reference's generator (test/_random_circuit_generator.py:61-86) minus
# barrier / measure / reset, which the JKU tests also strip (test_qasm_simulator_jku.py:57-62).
RANDOM_OPS = {
    "ccx": (3, 0), "ch": (2, 0), "crz": (2, 1), "cswap": (3, 0), "cu1": (2, 1), "cu3": (2, 3),
    "cx": (2, 0), "cy": (2, 0), "cz": (2, 0), "h": (1, 0), "iden": (1, 0), "rx": (1, 1),
    "ry": (1, 1), "rz": (1, 1), "s": (1, 0), "t": (1, 0), "u1": (1, 1), "u2": (1, 2),
    "u3": (1, 3), "x": (1, 0), "y": (1, 0), "z": (1, 0),
}


def random_circuit(n, n_ops, seed, basis=None):
    """BASELINE config 4: 'depth' = number of named ops (generator :159,222); parameters
    random.random() in [0,1) (:177); distinct qubits by random.sample (:219)."""
    rng = random.Random(seed)
    names = sorted(basis or RANDOM_OPS)
    names = [k for k in names if RANDOM_OPS[k][0] <= n]
    c = Circuit(n, name="random%d_%d_s%d" % (n, n_ops, seed))
    for _ in range(n_ops):
        name = rng.choice(names)
        nregs, nparams = RANDOM_OPS[name]
        params = [rng.random() for _ in range(nparams)]
        qubits = rng.sample(range(n), nregs)
        c.add(name, qubits, params)
    return c


def dense_statevector(circ):
    """Plain NumPy state-vector simulation with the same U convention; index bit q = qubit q.
    Independent check for small n (the stand-in for BasicAer in the reference's tests)."""
    import numpy as np

    n = circ.n
    psi = np.zeros(1 << n, dtype=np.complex128)
    psi[0] = 1.0
    idx = np.arange(1 << n)
    for op in circ.ops:
        m = op.matrix()
        t = op.target
        mask = np.ones(1 << n, dtype=bool)
        for q, v in op.controls:
            mask &= ((idx >> q) & 1) == v
        lo = mask & (((idx >> t) & 1) == 0)
        i0 = idx[lo]
        i1 = i0 | (1 << t)
        a0, a1 = psi[i0].copy(), psi[i1].copy()
        psi[i0] = m[0][0] * a0 + m[0][1] * a1
        psi[i1] = m[1][0] * a0 + m[1][1] * a1
    return psi


def teleport_qasm():
    """Teleportation of |1> from q[0] to q[2] with measured corrections: the smallest program with gates behind
    measurements, i.e. the per-shot re-simulation workload (QASMsimulator.cpp:501-522).  q[2] reads 1 in every shot."""
    return ('OPENQASM 2.0;\ninclude "qelib1.inc";\nqreg q[3];\ncreg c0[1];\ncreg c1[1];\ncreg out[1];\n'
            'x q[0];\nh q[1];\ncx q[1], q[2];\ncx q[0], q[1];\nh q[0];\nmeasure q[0] -> c0[0];\nmeasure q[1] -> c1[0];\n'
            'if (c0 == 1) z q[2];\nif (c1 == 1) x q[2];\nmeasure q[2] -> out[0];\n')

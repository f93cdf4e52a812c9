"""Thin Python wrapper over one ``jkq_handle`` of libjkq_b200.so (include/jkq.h)."""

import ctypes as C

import numpy as np

from . import _native
from .jkusimulatorerror import JKUSimulatorError

OP_GATE, OP_SNAPSHOT, OP_MEASURE, OP_BARRIER, OP_MEASURE_ONE, OP_RESET, OP_IF, OP_PROBABILITIES = 0, 1, 2, 3, 4, 5, 6, 7
FINAL_SLOT = 0xFFFFFFFF

SNODE_DTYPE = np.dtype([("c0", "<u4"), ("c1", "<u4"), ("q0", "<f8")])
# memory layout of jkq_op (include/jkq.h)
OP_DTYPE = np.dtype([("kind", "<u4"), ("target", "<u4"), ("ctrl_pos", "<u8"), ("ctrl_neg", "<u8"), ("m", "<f8", (8,)),
                     ("arg", "<u4"), ("reserved", "<u4")])
assert OP_DTYPE.itemsize == C.sizeof(_native.JkqOp)


def gate_op(target, matrix, controls=()):
    """jkq_op for a 2x2 ``matrix`` (nested sequence of complex) on ``target``; controls = ((qubit, value), ...)."""
    op = _native.JkqOp()
    op.kind = OP_GATE
    op.target = int(target)
    pos = neg = 0
    for q, v in controls:
        if v:
            pos |= 1 << int(q)
        else:
            neg |= 1 << int(q)
    op.ctrl_pos, op.ctrl_neg = pos, neg
    k = 0
    for i in range(2):
        for j in range(2):
            z = complex(matrix[i][j])
            op.m[k], op.m[k + 1] = z.real, z.imag
            k += 2
    return op


def circuit_ops(circuit, measure=None, final_snapshot=False):
    """circuits.Circuit -> numpy array of OP_DTYPE (one GATE per primitive, then the optional final snapshot and
    the measure map).  Cached on the circuit object for the plain case."""
    plain = not measure and not final_snapshot
    if plain and getattr(circuit, "_jkq_ops", None) is not None and len(circuit._jkq_ops) == len(circuit.ops):
        return circuit._jkq_ops
    n_extra = (1 if final_snapshot else 0) + len(measure or {})
    arr = np.zeros(len(circuit.ops) + n_extra, dtype=OP_DTYPE)
    mcache = {}
    for i, op in enumerate(circuit.ops):
        key = (op.kind, op.name, op.theta.const, op.theta.frac, op.phi.const, op.phi.frac, op.lam.const, op.lam.frac)
        m = mcache.get(key)
        if m is None:
            mm = op.matrix()
            m = mcache[key] = [x for row in mm for z in row for x in (complex(z).real, complex(z).imag)]
        pos = neg = 0
        for q, v in op.controls:
            if v:
                pos |= 1 << q
            else:
                neg |= 1 << q
        arr[i] = (OP_GATE, op.target, pos, neg, m, 0, 0)
    k = len(circuit.ops)
    if final_snapshot:
        arr[k] = (OP_SNAPSHOT, 0, 0, 0, [0.0] * 8, 0, 0)
        k += 1
    for q, cbit in (measure or {}).items():
        arr[k] = (OP_MEASURE, int(q), 0, 0, [0.0] * 8, int(cbit), 0)
        k += 1
    if plain:
        try:
            circuit._jkq_ops = arr
        except AttributeError:
            pass
    return arr


def simple_op(kind, target=0, arg=0):
    op = _native.JkqOp()
    op.kind, op.target, op.arg = kind, int(target), int(arg)
    return op


class Engine:
    """One engine instance bound to one CUDA device.  Not thread-safe."""

    def __init__(self, device=0, **cfg):
        self._lib = _native.load()
        c = _native.JkqConfig()
        c.device = int(device)
        for k, v in cfg.items():
            if not hasattr(c, k):
                raise TypeError("unknown engine option %r" % k)
            setattr(c, k, v)
        self._h = C.c_void_p()
        rc = self._lib.jkq_create(C.byref(self._h), C.byref(c))
        if rc != 0:
            msg = self._lib.jkq_last_error(self._h).decode() if self._h else "no usable CUDA device"
            if self._h:
                self._lib.jkq_destroy(self._h)
                self._h = None
            raise JKUSimulatorError("jkq_create failed (%d): %s" % (rc, msg))
        self.n = 0

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.jkq_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc != 0:
            raise JKUSimulatorError("jkq error %d: %s" % (rc, self._lib.jkq_last_error(self._h).decode()))

    # -- circuit ----------------------------------------------------------------------------
    def load_ops(self, n_qubits, ops):
        """``ops``: sequence of _native.JkqOp, or a numpy array of OP_DTYPE (passed without a copy)."""
        if isinstance(ops, np.ndarray):
            ops = np.ascontiguousarray(ops, dtype=OP_DTYPE)
            ptr = ops.ctypes.data_as(C.POINTER(_native.JkqOp)) if len(ops) else None
            self._check(self._lib.jkq_load_circuit(self._h, int(n_qubits), ptr, len(ops)))
        else:
            arr = (_native.JkqOp * max(1, len(ops)))(*ops)
            self._check(self._lib.jkq_load_circuit(self._h, int(n_qubits), arr, len(ops)))
        self.n = int(n_qubits)

    def load(self, circuit, measure=None, final_snapshot=False):
        """``circuit``: circuits.Circuit; ``measure``: {qubit: clbit} (None = all qubits, bit q = qubit q)."""
        self.load_ops(circuit.n, circuit_ops(circuit, measure, final_snapshot))

    def set_option(self, name, value):
        self._check(self._lib.jkq_set_option(self._h, name.encode(), int(value)))

    # -- execution --------------------------------------------------------------------------
    def apply(self):
        self._check(self._lib.jkq_apply(self._h))

    def sample(self, shots, seed=1, shot_offset=0, accumulate=False):
        self._check(self._lib.jkq_sample(self._h, int(shots), int(seed), int(shot_offset), int(bool(accumulate))))

    def run(self, shots, seed=1):
        self._check(self._lib.jkq_run(self._h, int(shots), int(seed)))

    # -- results ----------------------------------------------------------------------------
    def counts_arrays(self):
        n = C.c_size_t(0)
        self._check(self._lib.jkq_get_counts(self._h, None, None, 0, C.byref(n)))
        k = np.zeros(max(1, n.value), dtype=np.uint64)
        v = np.zeros(max(1, n.value), dtype=np.uint64)
        self._check(self._lib.jkq_get_counts(self._h, k.ctypes.data_as(C.POINTER(C.c_uint64)),
                                             v.ctypes.data_as(C.POINTER(C.c_uint64)), k.size, C.byref(n)))
        return k[:n.value], v[:n.value]

    def free_histogram(self):
        self._check(self._lib.jkq_release_histogram(self._h))

    def distinct_outcomes(self):
        """number of distinct outcomes in the histogram (no copy)"""
        n = C.c_size_t(0)
        self._check(self._lib.jkq_get_counts(self._h, None, None, 0, C.byref(n)))
        return int(n.value)

    def counts(self):
        """{outcome as int (bit c = classical bit c): count}"""
        k, v = self.counts_arrays()
        return {int(a): int(b) for a, b in zip(k, v)}

    def counts_bitstrings(self, width=None):
        """{string with character i = bit i: count} -- the layout the reference engine prints
        (QASMsimulator.cpp:491-494)."""
        width = width or self.n
        return {"".join("1" if (k >> i) & 1 else "0" for i in range(width)): v for k, v in self.counts().items()}

    def statevector(self, slot=FINAL_SLOT):
        out = np.zeros(2 << self.n, dtype=np.float64)
        self._check(self._lib.jkq_get_statevector(self._h, int(slot), out.ctypes.data_as(C.POINTER(C.c_double)), 1 << self.n))
        return out.view(np.complex128)

    def amplitudes(self, first, count):
        out = np.zeros(2 * max(1, count), dtype=np.float64)
        self._check(self._lib.jkq_get_amplitudes(self._h, int(first), int(count), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out.view(np.complex128)[:count]

    def snapshot_amplitudes(self, slot, first, count):
        """Amplitudes [first, first + count) of the state the SNAPSHOT op of ``slot`` recorded (``jkq_get_snapshot_amplitudes``):
        a slice of the dense vector, or -- above ``snapshot_dense_max`` (26) qubits, where the snapshot is a copy of the
        diagram -- one walk per amplitude.  Read a wide register range by range::

            for first in range(0, 1 << n, 1 << 20):
                chunk = eng.snapshot_amplitudes(slot, first, 1 << 20)"""
        out = np.zeros(2 * max(1, count), dtype=np.float64)
        self._check(self._lib.jkq_get_snapshot_amplitudes(self._h, int(slot), int(first), int(count), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out.view(np.complex128)[:count]

    def probabilities_range(self, slot, first, count):
        out = np.zeros(max(1, count), dtype=np.float64)
        self._check(self._lib.jkq_get_probabilities_range(self._h, int(slot), int(first), int(count), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out[:count]

    def level_counts(self):
        out = np.zeros(self.n, dtype=np.uint32)
        self._check(self._lib.jkq_get_level_counts(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), self.n))
        return [int(x) for x in out]

    def level_log(self, n_gates):
        out = np.zeros(max(1, n_gates * self.n), dtype=np.uint32)
        w = C.c_size_t(0)
        self._check(self._lib.jkq_get_level_log(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), out.size, C.byref(w)))
        return out[:w.value].reshape(n_gates, self.n).tolist()

    def probabilities(self, slot):
        """The 2^k probabilities a PROBABILITIES op recorded in ``slot`` (bit j of the index = j-th lowest listed qubit)."""
        n = C.c_size_t(0)
        self._check(self._lib.jkq_get_probabilities(self._h, int(slot), None, 0, C.byref(n)))
        out = np.zeros(n.value, dtype=np.float64)
        self._check(self._lib.jkq_get_probabilities(self._h, int(slot), out.ctypes.data_as(C.POINTER(C.c_double)), out.size, C.byref(n)))
        return out

    def qubit_probability(self, qubit):
        """P(qubit reads 0) in the current state (Simulator::AssignProbsOne)."""
        p = C.c_double(0)
        self._check(self._lib.jkq_get_qubit_probability(self._h, int(qubit), C.byref(p)))
        return p.value

    def mid_measurements(self):
        """{classical bit: outcome} of the MEASURE_ONE ops of the last apply."""
        bits, mask = C.c_uint64(0), C.c_uint64(0)
        self._check(self._lib.jkq_get_mid_measurements(self._h, C.byref(bits), C.byref(mask)))
        return {c: (bits.value >> c) & 1 for c in range(64) if (mask.value >> c) & 1}

    def stats(self):
        s = _native.JkqStats()
        self._check(self._lib.jkq_get_stats(self._h, C.byref(s)))
        return {name: getattr(s, name) for name, _ in s._fields_}

    # -- multi-GPU plumbing -----------------------------------------------------------------
    def dd_export(self):
        n, root = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.jkq_dd_export(self._h, None, 0, C.byref(n), C.byref(root)))
        arr = np.zeros(n.value, dtype=SNODE_DTYPE)
        self._check(self._lib.jkq_dd_export(self._h, arr.ctypes.data_as(C.POINTER(_native.JkqSNode)), arr.size,
                                            C.byref(n), C.byref(root)))
        return arr, int(root.value)

    def measure_map(self):
        out = np.zeros(max(1, self.n), dtype=np.uint8)
        self._check(self._lib.jkq_get_measure_map(self._h, out.ctypes.data_as(C.POINTER(C.c_uint8)), out.size))
        return out[:self.n]

    def dd_import(self, n_qubits, nodes, root, clbit_of_level):
        nodes = np.ascontiguousarray(nodes, dtype=SNODE_DTYPE)
        m = np.ascontiguousarray(clbit_of_level, dtype=np.uint8)
        self._check(self._lib.jkq_dd_import(self._h, int(n_qubits), nodes.ctypes.data_as(C.POINTER(_native.JkqSNode)),
                                            nodes.size, int(root), m.ctypes.data_as(C.POINTER(C.c_uint8))))
        self.n = int(n_qubits)

    def dd_size(self):
        """(nodes, root) of the sampling DD (builds it if need be)"""
        n, root = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.jkq_dd_export_device(self._h, None, 0, C.byref(n), C.byref(root)))
        return int(n.value), int(root.value)

    def dd_export_device(self, dev_ptr, cap_nodes):
        """Copy the sampling DD (16 bytes per node) into a DEVICE buffer of the caller (raw pointer): (nodes, root)."""
        n, root = C.c_uint32(0), C.c_uint32(0)
        self._check(self._lib.jkq_dd_export_device(self._h, C.c_void_p(dev_ptr), int(cap_nodes), C.byref(n), C.byref(root)))
        return int(n.value), int(root.value)

    def dd_import_device(self, n_qubits, dev_ptr, n_nodes, root, clbit_of_level):
        m = np.ascontiguousarray(clbit_of_level, dtype=np.uint8)
        self._check(self._lib.jkq_dd_import_device(self._h, int(n_qubits), C.c_void_p(dev_ptr), int(n_nodes), int(root),
                                                   m.ctypes.data_as(C.POINTER(C.c_uint8))))
        self.n = int(n_qubits)

    def sample_multi(self, devices, shots, seed=1):
        """jkq_sample_multi_gpu: the shots of the current state split over ``devices`` (CUDA ordinals of this process),
        sampling DD broadcast by peer copy, histograms merged on the device; ``counts()`` afterwards as after ``sample``.
        Returns the call's statistics."""
        dv = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        st = _native.JkqMultiStats()
        self._check(self._lib.jkq_sample_multi_gpu(self._h, dv, len(devices), int(shots), int(seed), C.byref(st)))
        return {name: getattr(st, name) for name, _ in st._fields_}

    def counts_partition(self, world, dev_keys=0, dev_counts=0, cap=0):
        """Sort the histogram by owner rank into device buffers (raw pointers; 0 = only count): per-owner sizes."""
        per = (C.c_uint64 * 64)()
        self._check(self._lib.jkq_counts_partition(self._h, int(world), C.c_void_p(dev_keys or None), C.c_void_p(dev_counts or None),
                                                   int(cap), per))
        return [int(per[r]) for r in range(world)]

    def counts_merge(self, dev_keys, dev_counts, n, expect=0, reset=True):
        self._check(self._lib.jkq_counts_merge(self._h, C.c_void_p(dev_keys or None), C.c_void_p(dev_counts or None), int(n), int(expect),
                                               int(bool(reset))))

    def random_sector_bench(self, nbytes=1 << 32, iters=3):
        a, b = C.c_double(0), C.c_double(0)
        self._check(self._lib.jkq_random_sector_bench(self._h, int(nbytes), int(iters), C.byref(a), C.byref(b)))
        return {"gsectors_per_s": a.value, "random_sector_gbs": a.value * 32.0, "gcas_per_s": b.value}

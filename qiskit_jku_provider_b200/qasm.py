"""OpenQASM 2.0 text -> engine ops through the native front end (``jkq_qasm_parse``, csrc/jkq_qasm.cpp).

The reference parses and executes QASM inside its executable (src/QASMsimulator.cpp:895-1060); here the same
grammar produces the ``jkq_op`` array that ``Engine.load_ops`` takes.  Host code only: no GPU is touched.
"""

import ctypes as C

import numpy as np

from . import _native
from .engine import OP_DTYPE
from .jkusimulatorerror import JKUSimulatorError

STATEVECTOR, PROBABILITIES = 1, 2  # JKQ_QASM_* flags of include/jkq.h


class QasmProgram:
    """Parsed program: ``ops`` (numpy array of OP_DTYPE), registers as (name, offset, size), snapshot statements as
    (slot, [qubits in argument order]), and what the driver needs to pick the shot strategy."""

    def __init__(self, n_qubits, n_clbits, ops, qregs, cregs, snapshots, gate_after_measure, first_measure):
        self.n_qubits, self.n_clbits, self.ops = n_qubits, n_clbits, ops
        self.qregs, self.cregs, self.snapshots = qregs, cregs, snapshots
        self.gate_after_measure, self.first_measure = gate_after_measure, first_measure


def parse_qasm(text, include_dir=None, statevector=False, probabilities=False):
    lib = _native.load()
    prog = C.c_void_p()
    flags = (STATEVECTOR if statevector else 0) | (PROBABILITIES if probabilities else 0)
    rc = lib.jkq_qasm_parse(text.encode(), include_dir.encode() if include_dir else None, flags, C.byref(prog))
    try:
        if rc != 0:
            raise JKUSimulatorError(lib.jkq_program_error(prog).decode() if prog else "jkq_qasm_parse failed")
        info = _native.JkqProgramInfo()
        lib.jkq_program_get_info(prog, C.byref(info))
        ops = np.zeros(info.n_ops, dtype=OP_DTYPE)
        if info.n_ops:
            C.memmove(ops.ctypes.data, info.ops, info.n_ops * OP_DTYPE.itemsize)
        regs = ([], [])
        for creg, count in ((0, info.n_qregs), (1, info.n_cregs)):
            for i in range(count):
                name = C.create_string_buffer(256)
                off, size = C.c_uint32(0), C.c_uint32(0)
                lib.jkq_program_get_register(prog, creg, i, name, 256, C.byref(off), C.byref(size))
                regs[creg].append((name.value.decode(), off.value, size.value))
        snaps = []
        for i in range(info.n_snapshots):
            slot, k = C.c_uint32(0), C.c_size_t(0)
            q = (C.c_uint32 * 64)()
            lib.jkq_program_get_snapshot(prog, i, C.byref(slot), q, 64, C.byref(k))
            snaps.append((slot.value, [int(q[j]) for j in range(k.value)]))
        return QasmProgram(info.n_qubits, info.n_clbits, ops, regs[0], regs[1], snaps, bool(info.gate_after_measure), int(info.first_measure))
    finally:
        if prog:
            lib.jkq_program_free(prog)

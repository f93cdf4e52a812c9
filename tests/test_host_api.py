"""CPU: host-side logic -- the C ABI library loads and exports every symbol include/jkq.h declares
(no compute without a GPU), the provider/backend/job mirror, Qobj conversion, count formatting."""
import numpy as np
import ctypes
import os
import re

import pytest

from conftest import ROOT
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200 import compat


def test_library_builds_and_exports_every_declared_symbol():
    from qiskit_jku_provider_b200 import _build, _native
    lib_path = _build.build_native()
    lib = ctypes.CDLL(lib_path)
    header = open(os.path.join(ROOT, "include", "jkq.h")).read()
    declared = set(re.findall(r"\b(jkq_[a-z_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == {n for n, _, _ in _native.SYMBOLS}


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from qiskit_jku_provider_b200.engine import Engine
    from qiskit_jku_provider_b200 import JKUSimulatorError
    with pytest.raises(JKUSimulatorError):
        Engine(device=0)


def test_qelib1_expansion_counts():
    # primitive counts of SURVEY 3.7
    sizes = {"cz": 3, "cy": 3, "swap": 3, "ch": 11, "ccx": 15, "cswap": 17, "crz": 4, "cu1": 5, "cu3": 5, "rzz": 3,
             "h": 1, "x": 1, "u0": 1, "iden": 1}
    for name, k in sizes.items():
        nq, npar = compat._GATES.get(name, (None, None)) if name in compat._GATES else (C.RANDOM_OPS[name][0], C.RANDOM_OPS[name][1])
        ops = C.expand(name, list(range(nq)), [0.3] * npar)
        assert len(ops) == k, name
    with pytest.raises(KeyError):
        C.expand("nosuchgate", [0])


def test_u_matrix_convention():
    import numpy as np
    # x = U(pi,0,pi) = -iX ; u1(l) = diag(e^{-il/2}, e^{il/2})  (QASMsimulator.cpp:295-298)
    x = np.array(C.expand("x", [0])[0].matrix())
    assert np.allclose(x, -1j * np.array([[0, 1], [1, 0]]), atol=1e-15)
    u1 = np.array(C.expand("u1", [0], [0.7])[0].matrix())
    assert np.allclose(u1, np.diag([np.exp(-0.35j), np.exp(0.35j)]), atol=1e-15)


def _bell_qobj(measure=True):
    qr, cr = compat.QuantumRegister(2, "q"), compat.ClassicalRegister(2, "c")
    qc = compat.QuantumCircuit(qr, cr, name="bell")
    qc.h(qr[0])
    qc.cx(qr[0], qr[1])
    if measure:
        qc.measure(qr, cr)
    return compat.assemble(qc, shots=100, seed=3)


def test_qobj_conversion_without_gpu():
    from qiskit_jku_provider_b200.qasm_simulator_jku import JKUSimulatorWrapper
    w = JKUSimulatorWrapper(engine=None, silent=True)
    exp = _bell_qobj().experiments[0]
    md = w.compute_measurement_data(exp)
    assert md["mapping"] == {0: 0, 1: 1} and md["clbits_num"] == 2 and md["qubits_num"] == 2
    ops, slots = w.convert_qobj_circuit_to_ops(exp, True)
    assert ops["kind"].tolist() == [0, 0, 2, 2, 1]  # h, cx, measure x2, final snapshot
    assert slots == [1] and w.max_snapshot_index == 1
    assert ops[1]["ctrl_pos"] == 1 and ops[1]["target"] == 1
    assert ops.dtype.itemsize == 96  # sizeof(jkq_op), include/jkq.h
    h = 2 ** -0.5
    assert np.allclose(ops[0]["m"], [h, 0, h, 0, h, 0, -h, 0]) or np.allclose(np.abs(ops[0]["m"][::2] + 1j * ops[0]["m"][1::2]), [h] * 4)


def test_qubit_reuse_after_measure_raises():
    from qiskit_jku_provider_b200 import JKUSimulatorError
    from qiskit_jku_provider_b200.qasm_simulator_jku import JKUSimulatorWrapper
    qr, cr = compat.QuantumRegister(1, "q"), compat.ClassicalRegister(1, "c")
    qc = compat.QuantumCircuit(qr, cr)
    qc.measure(qr[0], cr[0])
    qc.x(qr[0])
    with pytest.raises(JKUSimulatorError):
        JKUSimulatorWrapper(None, True).compute_measurement_data(compat.assemble(qc).experiments[0])


def test_count_formatting_with_registers():
    # '01 10' style of test/test_multi_registers_convention.py:45
    out = compat._format_counts({"0x6": 1024}, [["c0", 2], ["c1", 2]], 4)
    assert out == {"01 10": 1024}
    assert compat._format_counts({"0x1": 5}, [["c", 1]], 1) == {"1": 5}


def test_job_requires_submit_and_rejects_double_submit():
    from qiskit_jku_provider_b200.jkujob import JKUJob
    job = JKUJob(None, "id", lambda job_id, qobj: 42, _bell_qobj())
    with pytest.raises(compat.JobError):
        job.result()
    job.submit()
    assert job.result(timeout=10) == 42
    assert job.status() == compat.JobStatus.DONE
    with pytest.raises(compat.JobError):
        job.submit()


def test_drop_in_import_name():
    import qiskit_jku_provider
    assert qiskit_jku_provider.JKUSimulatorError is not None
    assert qiskit_jku_provider.QasmSimulator.DEFAULT_CONFIGURATION["backend_name"] == "qasm_simulator"


def test_shard_ranges_cover_all_shots():
    from qiskit_jku_provider_b200.multi_gpu import shard
    for total in (0, 1, 7, 1024, 10 ** 6 + 3):
        for world in (1, 2, 3, 8):
            r = [shard(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1

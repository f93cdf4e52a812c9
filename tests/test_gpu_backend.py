"""GPU: the provider / backend / job API, restating the reference's own tests (test/*.py) without
qiskit (qiskit_jku_provider_b200.compat supplies the circuit builder and Result)."""
import numpy as np
import pytest

from qiskit_jku_provider_b200.compat import (ClassicalRegister, JobStatus, QuantumCircuit, QuantumRegister, assemble,
                                              execute)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend():
    from qiskit_jku_provider_b200 import QasmSimulator
    return QasmSimulator(silent=True, node_capacity=1 << 18)


def test_gate_x(backend):
    # test/test_qasm_simulator_jku.py:79-89
    qr, cr = QuantumRegister(1), ClassicalRegister(1)
    qc = QuantumCircuit(qr, cr, name='test_gate_x')
    qc.x(qr[0])
    qc.measure(qr, cr)
    assert execute(qc, backend, shots=100).result(timeout=30).get_counts() == {'1': 100}


def test_entangle(backend):
    # test/test_qasm_simulator_jku.py:91-109
    n = 5
    qr, cr = QuantumRegister(n), ClassicalRegister(n)
    qc = QuantumCircuit(qr, cr, name='test_entangle')
    qc.h(qr[0])
    for i in range(1, n):
        qc.cx(qr[0], qr[i])
    qc.measure(qr, cr)
    counts = execute(qc, backend, shots=100).result(timeout=30).get_counts()
    assert set(counts) <= {'0' * n, '1' * n} and sum(counts.values()) == 100


def test_output_style(backend):
    # test/test_qasm_simulator_jku.py:111-136 -- the chi-squared comparison against BasicAer becomes a
    # comparison against the exact distribution (h q0: outcomes '00' and '01' with 1/2 each)
    from scipy.stats import chisquare
    qr = QuantumRegister(2)
    cr = ClassicalRegister(2)
    qc = QuantumCircuit(qr, cr, name='test_output_order')
    qc.h(qr[0])
    qc.measure(qr[0], cr[0])
    qc.measure(qr[1], cr[1])
    counts = execute(qc, backend, shots=1000, seed=1).result().get_counts()
    assert set(counts) == {'00', '01'}
    assert chisquare([counts['00'], counts['01']]).pvalue > 0.01
    cr1, cr2 = ClassicalRegister(1), ClassicalRegister(1)
    qc = QuantumCircuit(qr, cr1, cr2, name='test_output_separation')
    qc.h(qr[0])
    qc.measure(qr[0], cr1[0])
    qc.measure(qr[1], cr2[0])
    counts = execute(qc, backend, shots=1000, seed=1).result().get_counts()
    assert set(counts) == {'0 0', '0 1'}


def test_random_circuits_against_dense_simulation(backend):
    # test/test_qasm_simulator_jku.py:138-158, BasicAer replaced by a NumPy state-vector simulation
    import random
    from scipy.stats import chisquare
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.compat import _GATES
    rng = random.Random(7)
    names = [k for k in C.RANDOM_OPS if k in _GATES]
    for trial in range(20):
        n = rng.randint(1, 4)
        depth = rng.randint(1, 50)
        basis = [g for g in rng.sample(names, rng.randint(2, 7)) if C.RANDOM_OPS[g][0] <= n] or ['h']
        qr, cr = QuantumRegister(n), ClassicalRegister(n)
        qc = QuantumCircuit(qr, cr, name='rand%d' % trial)
        prim = C.Circuit(n)
        for _ in range(depth):
            g = rng.choice(basis)
            nq, npar = C.RANDOM_OPS[g]
            params = [rng.random() for _ in range(npar)]
            qubits = rng.sample(range(n), nq)
            getattr(qc, g)(*(params + [qr[q] for q in qubits]))
            prim.add(g, qubits, params)
        qc.measure(qr, cr)
        shots = 2000
        res = execute(qc, backend, shots=shots, seed=trial + 1).result()
        counts = res.get_counts()
        probs = np.abs(C.dense_statevector(prim)) ** 2
        assert sum(counts.values()) == shots
        obs, exp = [], []
        rest_o = rest_e = 0.0
        for i, p in enumerate(probs):
            o = counts.get(format(i, '0%db' % n), 0)
            if p * shots >= 5:
                obs.append(o)
                exp.append(p * shots)
            else:
                rest_o += o
                rest_e += p * shots
        if rest_e > 1e-9:
            obs.append(rest_o)
            exp.append(rest_e)
        if len(obs) > 1:
            exp = np.array(exp) * (sum(obs) / sum(exp))
            assert chisquare(obs, exp).pvalue > 1e-4, qc.name
        # final state vector equals the dense simulation (up to nothing: same U convention)
        assert np.abs(res.get_statevector() - C.dense_statevector(prim)).max() < 1e-9


def test_statevector_output(backend):
    # test/test_jku_snapshot.py:29-39
    qr = QuantumRegister(2)
    qc = QuantumCircuit(qr)
    qc.h(qr[0])
    qc.cx(qr[0], qr[1])
    result = execute(qc, backend).result()
    assert result.success is True
    actual = result.get_statevector(qc)
    assert abs(abs(actual[0]) ** 2 - 0.5) < 1e-5
    assert actual[1] == 0 and actual[2] == 0
    assert abs(abs(actual[3]) ** 2 - 0.5) < 1e-5


def test_circuit_multi_registers(backend):
    # test/test_multi_registers_convention.py:23-51
    qreg0, creg0 = QuantumRegister(2, 'q0'), ClassicalRegister(2, 'c0')
    qreg1, creg1 = QuantumRegister(2, 'q1'), ClassicalRegister(2, 'c1')
    circ = QuantumCircuit(qreg0, qreg1)
    circ.x(qreg0[1])
    circ.x(qreg1[0])
    meas = QuantumCircuit(qreg0, qreg1, creg0, creg1)
    meas.measure(qreg0, creg0)
    meas.measure(qreg1, creg1)
    qc = circ + meas
    counts = execute(qc, backend).result().get_counts(qc)
    assert counts == {'01 10': 1024}
    state = execute(circ, backend).result().get_statevector(circ)
    basis = np.zeros(16, dtype=complex)
    basis[int('0110', 2)] = 1
    assert abs(abs(np.vdot(basis, state)) ** 2 - 1.0) < 1e-7


def test_example_qasm(backend):
    # test/test_qasm_simulator.py:29-44 (qasms/example.qasm: h q; cx q,r; measure both registers)
    q, r = QuantumRegister(3, 'q'), QuantumRegister(3, 'r')
    c, d = ClassicalRegister(3, 'c'), ClassicalRegister(3, 'd')
    qc = QuantumCircuit(q, r, c, d, name='example')
    qc.h(q)
    qc.cx(q, r)
    qc.barrier(q)
    qc.measure(q, c)
    qc.measure(r, d)
    assert execute(qc, backend, shots=1).result().success
    shots = 1024
    counts = execute(qc, backend, shots=shots, seed=1).result().get_counts()
    expected = ['000 000', '001 001', '010 010', '011 011', '100 100', '101 101', '110 110', '111 111']
    assert sorted(counts) == expected
    for v in counts.values():
        assert abs(v - shots / 8) <= shots * 0.04


def test_backend_and_provider_api(backend):
    # test/test_jku_backend.py:24-43, test/test_jku_provider.py:22-30, examples/*.py
    from qiskit_jku_provider_b200 import JKUProvider
    provider = JKUProvider()
    assert str(provider) == 'JKUProvider'
    assert len(provider.backends()) > 0
    b = provider.get_backend('qasm_simulator')
    assert b.name() == 'qasm_simulator'
    assert provider.available_backends({'local': True}) == [b]
    assert provider.available_backends({'local': False}) == []
    assert b.configuration().backend_name == 'qasm_simulator'
    assert b.properties() is None
    assert b.status().operational is True
    qr, cr = QuantumRegister(2), ClassicalRegister(2)
    qc = QuantumCircuit(qr, cr)
    qc.h(qr[0])
    qc.cx(qr[0], qr[1])
    qc.measure(qr, cr)
    qobj = assemble(qc, shots=64, seed=2)
    job = b.run(qobj)
    res = job.result(timeout=60)
    assert job.status() == JobStatus.DONE and job.backend() is b and job.qobj() is qobj
    assert res.success and res.status == 'COMPLETED' and res.qobj_id == qobj.qobj_id
    r0 = res.results[0]
    assert r0.shots == 64 and r0.seed == 2 and r0.success and r0.status == 'DONE'
    assert set(r0.data['counts']) <= {'0x0', '0x3'}


def test_probabilities_output(backend):
    qr = QuantumRegister(2)
    qc = QuantumCircuit(qr)
    qc.h(qr[0])
    qc.snapshot(3)
    qc.cx(qr[0], qr[1])
    res = execute(qc, backend, data=['probabilities', 'probabilities_ket']).result()
    snaps = res.data()['snapshots']
    assert set(snaps) == {'3', '4'}
    assert np.allclose(snaps['3']['probabilities'], [0.5, 0.5, 0, 0])
    assert set(snaps['4']['probabilities_ket']) == {'00', '11'}
    assert res.data()['counts'] == {}


def test_reset_projects_like_the_reference():
    """h q0; cx q0,q1; reset q0; measure: the reference's ResetQubit projects on q0 = 0 (Simulator.cpp:291-339),
    which leaves |00> -- every shot reads 00."""
    from qiskit_jku_provider_b200 import JKUProvider
    from qiskit_jku_provider_b200.compat import ClassicalRegister, QuantumCircuit, QuantumRegister, assemble
    qr, cr = QuantumRegister(2, "q"), ClassicalRegister(2, "c")
    qc = QuantumCircuit(qr, cr, name="reset")
    qc.h(qr[0])
    qc.cx(qr[0], qr[1])
    qc.reset(qr[0])
    qc.measure(qr, cr)
    backend = JKUProvider().get_backend("qasm_simulator")
    backend.silent = True
    assert backend.run(assemble(qc, shots=300, seed=5)).result().get_counts() == {"00": 300}


def test_physical_reset_is_resimulated_per_shot():
    """With physical_reset: q0 always reads 0, q1 keeps the branch the measurement collapsed it to -- a 50/50 mixture
    over shots (re-simulated per shot, as the reference does for non-reusable states, QASMsimulator.cpp:501-522)."""
    from qiskit_jku_provider_b200 import JKUProvider
    from qiskit_jku_provider_b200.compat import ClassicalRegister, QuantumCircuit, QuantumRegister, assemble
    qr, cr = QuantumRegister(2, "q"), ClassicalRegister(2, "c")
    qc = QuantumCircuit(qr, cr, name="reset")
    qc.h(qr[0])
    qc.cx(qr[0], qr[1])
    qc.reset(qr[0])
    qc.measure(qr, cr)
    backend = JKUProvider().get_backend("qasm_simulator")
    backend.silent = True
    backend.physical_reset = True
    try:
        counts = backend.run(assemble(qc, shots=300, seed=5)).result().get_counts()
    finally:
        backend.physical_reset = False
    assert set(counts) <= {"00", "10"} and sum(counts.values()) == 300
    assert 100 <= counts.get("00", 0) <= 200


def _random_qc(seed):
    import random
    rnd = random.Random(seed)
    n = rnd.randint(3, 9)
    qr, cr = QuantumRegister(n, "q"), ClassicalRegister(n, "c")
    qc = QuantumCircuit(qr, cr, name="batch%d" % seed)
    for _ in range(rnd.randint(10, 60)):
        g = rnd.choice(["h", "x", "t", "s", "u3", "u1", "cx", "cz", "ccx", "cu1", "swap"])
        q = rnd.sample(range(n), 3)
        a = [rnd.uniform(-3, 3) for _ in range(3)]
        if g in ("h", "x", "t", "s"):
            getattr(qc, g)(qr[q[0]])
        elif g == "u1":
            qc.u1(a[0], qr[q[0]])
        elif g == "u3":
            qc.u3(a[0], a[1], a[2], qr[q[0]])
        elif g in ("cx", "cz", "swap"):
            getattr(qc, g)(qr[q[0]], qr[q[1]])
        elif g == "ccx":
            qc.ccx(qr[q[0]], qr[q[1]], qr[q[2]])
        else:
            qc.cu1(a[0], qr[q[0]], qr[q[1]])
    if seed % 3 == 0:  # a partial, permuted measure map
        for k in range(0, n, 2):
            qc.measure(qr[k], cr[n - 1 - k])
    else:
        qc.measure(qr, cr)
    return qc


def test_count_only_qobj_runs_as_one_batch_with_the_counts_of_the_sequential_path():
    """A Qobj of many experiments that return counts only (statevector=False) goes through jkq_batch_run -- one
    persistent kernel, one circuit per SM -- where the reference loops over one process per experiment
    (qasm_simulator_jku.py:340-359); the counts are those of the one-after-the-other path, experiment by experiment."""
    from qiskit_jku_provider_b200 import QasmSimulator
    qcs = [_random_qc(s) for s in range(40)]
    qobj = assemble(qcs, shots=500, seed=11)
    seq = QasmSimulator(silent=True, node_capacity=1 << 18)
    bat = QasmSimulator(silent=True, node_capacity=1 << 18, statevector=False)
    assert seq._batch_plan(qobj) is None and bat._batch_plan(qobj) is not None
    r_seq, r_bat = seq.run(qobj).result(), bat.run(qobj).result()
    for i, qc in enumerate(qcs):
        assert r_bat.get_counts(qc.name) == r_seq.get_counts(qc.name), qc.name
        assert 'statevector' not in r_bat.data(i) and 'statevector' in r_seq.data(i)
    assert 'batch' in r_bat.to_dict()['results'][0]['metadata']
    # an experiment that asks for its state vector takes the whole Qobj back to the sequential path
    qobj_sv = assemble(qcs[:5], shots=100, seed=3, data=['statevector'])
    assert bat._batch_plan(qobj_sv) is None
    assert len(bat.run(qobj_sv).result().get_statevector(qcs[0].name)) == 1 << qcs[0].qregs[0].size

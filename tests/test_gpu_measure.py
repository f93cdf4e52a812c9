"""GPU: mid-circuit measurement and reset (Simulator::MeasureOne / ResetQubit, Simulator.cpp:226-413) and the
single-qubit marginal (AssignProbsOne / GetProbability), through the C ABI."""
import numpy as np
import pytest

from qiskit_jku_provider_b200 import circuits as C

pytestmark = pytest.mark.gpu


def _ops(circ):
    from qiskit_jku_provider_b200.engine import gate_op
    return [gate_op(op.target, op.matrix(), op.controls) for op in circ.ops]


@pytest.fixture(scope="module")
def eng():
    from qiskit_jku_provider_b200.engine import Engine
    e = Engine(device=0, node_capacity=1 << 18)
    yield e
    e.close()


def test_qubit_probability_matches_the_state_vector(eng):
    for circ in (C.random_circuit(7, 40, 3), C.qft(6), C.ghz(5)):
        eng.load(circ)
        eng.apply()
        amp = eng.statevector()
        p = np.abs(amp) ** 2
        idx = np.arange(len(p))
        for q in range(circ.n):
            want = p[((idx >> q) & 1) == 0].sum()
            assert abs(eng.qubit_probability(q) - want) < 1e-9, (circ.name, q)


def test_measure_one_collapses_a_ghz_state(eng):
    from qiskit_jku_provider_b200.engine import OP_MEASURE_ONE, simple_op
    n = 8
    ones = 0
    for seed in range(40):
        ops = _ops(C.ghz(n)) + [simple_op(OP_MEASURE_ONE, 3, 0)]
        eng.set_option("measure_seed", 1000 + seed)
        eng.load_ops(n, ops)
        eng.run(256, seed)
        bit = eng.mid_measurements()[0]
        ones += bit
        assert eng.counts() == {((1 << n) - 1) * bit: 256}          # every shot reads the collapsed branch
        amp = eng.statevector()
        assert abs(abs(amp[((1 << n) - 1) * bit]) - 1.0) < 1e-9      # rescaled by 1/sqrt(p)
        assert abs(eng.stats()["norm"] - 1.0) < 1e-9
    assert 8 <= ones <= 32                                           # p = 1/2: 40 draws


def test_measure_then_gates_continue_on_the_collapsed_state(eng):
    from qiskit_jku_provider_b200.engine import OP_MEASURE_ONE, gate_op, simple_op
    h = C.Circuit(1).add("h", [0]).ops[0]
    ops = _ops(C.ghz(4)) + [simple_op(OP_MEASURE_ONE, 0, 5), gate_op(1, h.matrix(), ())]
    eng.set_option("measure_seed", 7)
    eng.load_ops(4, ops)
    eng.apply()
    b = eng.mid_measurements()[5]
    p = np.abs(eng.statevector()) ** 2
    base = 0b1101 * b          # qubits 0, 2, 3 keep the measured value, qubit 1 is in (|0> +- |1>)/sqrt2
    assert abs(p[base] - 0.5) < 1e-9 and abs(p[base | 2] - 0.5) < 1e-9 and abs(p.sum() - 1.0) < 1e-9


def test_reset_projects_on_zero_like_the_reference(eng):
    """Simulator::ResetQubit (Simulator.cpp:291-339): projection on |0> of the qubit, rescaled; X first when P(0) = 0."""
    from qiskit_jku_provider_b200.engine import OP_RESET, simple_op
    eng.set_option("physical_reset", 0)
    eng.load_ops(3, _ops(C.ghz(3)) + [simple_op(OP_RESET, 0, 0)])
    eng.apply()
    p = np.abs(eng.statevector()) ** 2
    assert abs(p[0b000] - 1.0) < 1e-9 and abs(eng.qubit_probability(0) - 1.0) < 1e-12
    # |1> on the qubit: P(0) = 0, the reference flips it
    c = C.Circuit(2)
    c.add("x", [0])
    c.add("h", [1])
    eng.load_ops(2, _ops(c) + [simple_op(OP_RESET, 0, 0)])
    eng.apply()
    p = np.abs(eng.statevector()) ** 2
    assert abs(p[0b00] - 0.5) < 1e-9 and abs(p[0b10] - 0.5) < 1e-9
    assert eng.stats()["gates"] == 2  # the projector / X slots are not ApplyGate calls


def test_physical_reset_measures_and_flips(eng):
    from qiskit_jku_provider_b200.engine import OP_RESET, simple_op
    eng.set_option("physical_reset", 1)
    seen = set()
    try:
        for seed in range(12):
            ops = _ops(C.ghz(3)) + [simple_op(OP_RESET, 0, 0)]
            eng.set_option("measure_seed", 50 + seed)
            eng.load_ops(3, ops)
            eng.apply()
            p = np.abs(eng.statevector()) ** 2
            assert abs(eng.qubit_probability(0) - 1.0) < 1e-12           # qubit 0 is |0> again
            assert min(abs(p[0b000] - 1.0), abs(p[0b110] - 1.0)) < 1e-9  # the other two keep the collapsed branch
            seen.add(int(np.argmax(p)))
    finally:
        eng.set_option("physical_reset", 0)
    assert seen == {0b000, 0b110}


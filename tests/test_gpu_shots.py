"""GPU: per-shot re-simulation as one kernel (jkq_batch_run_shots / k_batch_shots) against the sequential host-driven
path (jkq_apply + jkq_sample per shot): mid-circuit MeasureOne / ResetQubit / `if` handled on the device must give,
shot for shot, the very outcomes the host path draws from the same Philox streams."""
import numpy as np
import pytest

from test_qasm_frontend import HEADER, load_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def runner():
    from qiskit_jku_provider_b200.batch import BatchRunner
    with BatchRunner(workers=148, device=0) as br:
        yield br


@pytest.fixture(scope="module")
def eng():
    from qiskit_jku_provider_b200.engine import Engine
    e = Engine(device=0, node_capacity=1 << 18)
    yield e
    e.close()


def sequential(eng, prog, shots, seed, measure_seed0, physical_reset=False):
    keys, bits = [], []
    eng.set_option("physical_reset", int(physical_reset))
    try:
        eng.load_ops(prog.n_qubits, prog.ops)
        for s in range(shots):
            eng.set_option("measure_seed", (measure_seed0 + s) & 0x7FFFFFFFFFFFFFFF)
            eng.apply()
            eng.sample(1, seed=seed, shot_offset=s)
            (k,), (c,) = (list(x) for x in eng.counts_arrays())
            assert c == 1
            keys.append(int(k))
            mm = eng.mid_measurements()
            bits.append(sum(v << b for b, v in mm.items()))
    finally:
        eng.set_option("physical_reset", 0)
    return keys, bits


PROGRAMS = {
    "teleport": None,
    "if_multi": HEADER + "qreg q[4]; creg c[2]; creg d[1];\nh q[0];\nh q[1];\nmeasure q[0] -> c[0];\nmeasure q[1] -> c[1];\n"
                         "if (c == 2) ccx q[0], q[1], q[2];\nif (c == 3) ccx q[0], q[1], q[2];\nif (c == 1) swap q[0], q[3];\n"
                         "ry(0.8) q[2];\nif (c == 0) measure q[2] -> d[0];\nif (d == 1) x q[3];\nh q[1];\n",
    "reset_chain": HEADER + "qreg q[3];\nry(1.1) q[0];\ncx q[0], q[1];\nry(0.4) q[2];\ncx q[2], q[0];\nreset q[0];\nx q[1];\nreset q[1];\n"
                            "h q[2];\ncreg c[1];\nmeasure q[2] -> c[0];\nif (c == 1) h q[0];\n",
    "ghz_feedback": HEADER + "qreg q[6]; creg c[6];\nh q[0];\ncx q[0],q[1];\ncx q[1],q[2];\ncx q[2],q[3];\ncx q[3],q[4];\ncx q[4],q[5];\n"
                             "measure q[2] -> c[2];\nif (c == 4) x q;\nh q[5];\nmeasure q[5] -> c[5];\nry(0.3) q[0];\n",
}


@pytest.mark.parametrize("name", sorted(PROGRAMS))
@pytest.mark.parametrize("physical", [False, True])
def test_shots_kernel_equals_the_sequential_path(runner, eng, name, physical):
    from qiskit_jku_provider_b200.qasm import parse_qasm
    text = PROGRAMS[name] or load_case(name)[0]
    prog = parse_qasm(text)
    shots, seed, ms0 = 96, 5, 1234567
    want_keys, want_bits = sequential(eng, prog, shots, seed, ms0, physical)
    keys, bits, status, gates, ms = runner.run_shots(prog.n_qubits, prog.ops, shots, seed=seed, measure_seed0=ms0, physical_reset=physical)
    assert not status.any()
    assert [int(k) for k in keys] == want_keys
    assert [int(b) for b in bits] == want_bits
    assert ms > 0 and gates > 0


def test_teleport_many_shots(runner):
    """20 000 shots of the teleportation circuit in one launch: q[2] always reads 1, (q0, q1) uniform"""
    from qiskit_jku_provider_b200.qasm import parse_qasm
    prog = parse_qasm(load_case("teleport")[0])
    keys, bits, status, gates, ms = runner.run_shots(prog.n_qubits, prog.ops, 20000, seed=3)
    assert not status.any()
    assert ((keys >> np.uint64(2)) & np.uint64(1)).all()
    hist = np.bincount((keys & np.uint64(3)).astype(np.int64), minlength=4)
    assert (np.abs(hist - 5000) < 300).all(), hist
    # the classical bits are what the corrections were conditioned on; out[0] (bit 2) is the teleported 1
    assert ((bits >> np.uint64(2)) & np.uint64(1)).all()
    assert 4.5 * 20000 < gates < 7.5 * 20000


def test_shot_offset_and_split_give_the_same_shots(runner):
    from qiskit_jku_provider_b200.qasm import parse_qasm
    prog = parse_qasm(PROGRAMS["if_multi"])
    full, _, st, _, _ = runner.run_shots(prog.n_qubits, prog.ops, 200, seed=9, measure_seed0=77)
    a, _, _, _, _ = runner.run_shots(prog.n_qubits, prog.ops, 120, seed=9, measure_seed0=77)
    b, _, _, _, _ = runner.run_shots(prog.n_qubits, prog.ops, 80, seed=9, shot_offset=120, measure_seed0=77 + 120)
    assert not st.any() and list(full) == list(a) + list(b)


def test_circuit_without_measurements_is_just_sampled(runner, eng):
    from qiskit_jku_provider_b200 import circuits as C
    from qiskit_jku_provider_b200.engine import circuit_ops
    circ = C.random_circuit(6, 30, 2)
    keys, _, status, _, _ = runner.run_shots(circ.n, circuit_ops(circ), 500, seed=4)
    eng.load(circ)
    eng.run(500, 4)
    want = eng.counts()
    got = {}
    for k in keys:
        got[int(k)] = got.get(int(k), 0) + 1
    assert not status.any() and got == want


def test_shots_sweep_subset():
    """scripts/shots_sweep.py on a slice of its seeds: random programs with mid-circuit measure / reset / if -- the
    per-shot kernel equals the sequential path shot for shot, and the final state of a shot is one of the trajectories
    of an independent dense NumPy simulation (1e-9)."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "shots_sweep.py"), "10", "100"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "shots sweep ok" in out.stdout

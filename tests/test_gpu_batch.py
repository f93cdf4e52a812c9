"""GPU: batches of independent circuits run side by side on different SMs (BASELINE configs[4]); results must
equal the one-at-a-time results bit for bit (same engine code, same Philox streams)."""
import pytest

from qiskit_jku_provider_b200 import circuits as C

pytestmark = pytest.mark.gpu


def test_batch_equals_sequential():
    from qiskit_jku_provider_b200.batch import BatchRunner
    from qiskit_jku_provider_b200.engine import Engine
    circs = [C.random_circuit(12, 40, s) for s in range(40)] + [C.qft(10), C.ghz(16), C.grover(8, 3, marked=0x55)]
    with BatchRunner(workers=16) as br:
        got, gates, _ = br.run(circs, shots=512, seed=100)
        again, _, _ = br.run(circs, shots=512, seed=100)   # engines are reused
    assert got == again
    assert gates == sum(len(c) for c in circs)
    with Engine(device=0, node_capacity=1 << 18) as eng:
        for i, c in enumerate(circs):
            eng.load(c)
            eng.run(512, 100 + i)
            assert eng.counts() == got[i], c.name


def test_batch_sharding_by_rank_covers_every_circuit_once():
    from qiskit_jku_provider_b200.batch import BatchRunner
    circs = [C.random_circuit(8, 30, s) for s in range(10)]
    with BatchRunner(workers=4) as br:
        whole, _, _ = br.run(circs, shots=64, seed=3)
        parts = [br.run(circs, shots=64, seed=3, rank=r, world=3)[0] for r in range(3)]
    merged = {}
    for p in parts:
        assert not set(p) & set(merged)
        merged.update(p)
    assert merged == whole


def test_batch_falls_back_to_the_big_engine():
    from qiskit_jku_provider_b200.batch import BatchRunner
    from qiskit_jku_provider_b200.engine import Engine
    circs = [C.random_circuit(14, 80, 2), C.ghz(6)]
    with BatchRunner(workers=2, small=dict(node_capacity=1 << 10, item_capacity=1 << 10), big=dict(node_capacity=1 << 20)) as br:
        got, _, _ = br.run(circs, shots=256, seed=9)
        assert br.fallbacks >= 1
    with Engine(device=0, node_capacity=1 << 20) as eng:
        for i, c in enumerate(circs):
            eng.load(c)
            eng.run(256, 9 + i)
            assert eng.counts() == got[i]

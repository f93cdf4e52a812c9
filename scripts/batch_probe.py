"""Throughput of a batch of independent circuits vs the number of concurrent engines."""
import json, sys, time
sys.path.insert(0, ".")
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.batch import BatchRunner
from qiskit_jku_provider_b200.engine import circuit_ops

n_circ = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
circs = [C.random_circuit(18, 40, s) for s in range(n_circ)]
for c in circs:
    circuit_ops(c)
total = sum(len(c) for c in circs)
for w in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["1", "16", "64", "148"])]:
    with BatchRunner(workers=w) as br:
        br.run(circs, shots=1024, seed=1)  # warm-up: sizes every buffer
        t = time.perf_counter()
        out, gates, dev_ms = br.run(circs, shots=1024, seed=1)
        dt = time.perf_counter() - t
    print(json.dumps({"workers": w, "circuits": n_circ, "gate_applies": gates, "wall_s": round(dt, 4), "gate_applies_per_s": round(gates / dt),
                      "circuits_per_s": round(n_circ / dt, 1), "sum_device_ms": round(dev_ms, 1), "native_s": round(br.native_s, 4), "fallbacks": br.fallbacks}), flush=True)

"""Workloads for the round-2 ncu captures (one kernel family per mode).
usage: profile_r2.py sample_ghz | sample_qft | batch | shots | grid160 | qft32 | grover"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.engine import Engine

mode = sys.argv[1]
if mode in ("sample_ghz", "sample_qft"):
    circ, shots = (C.ghz(24), 10 ** 9) if mode == "sample_ghz" else (C.qft(32), 10 ** 8)
    with Engine(device=0, node_capacity=1 << 18) as eng:
        eng.load(circ)
        eng.apply()
        eng.sample(shots, seed=7)
        st = eng.stats()
        print(mode, "kernel_ms", st["sample_kernel_ms"], "shots/s", shots / st["sample_kernel_ms"] * 1e3, "distinct", eng.distinct_outcomes())
elif mode == "batch":
    from qiskit_jku_provider_b200.batch import BatchRunner
    circs = [C.random_circuit(18, 40, s) for s in range(1024)]
    with BatchRunner(workers=148) as br:
        _, gates, ms = br.run(circs, shots=1024, seed=1)
        print("batch gates", gates, "k_batch_ms", br.last_kernel_ms, "fallback_ms", br.last_fallback_ms)
elif mode == "shots":
    from qiskit_jku_provider_b200.batch import BatchRunner
    from qiskit_jku_provider_b200.qasm import parse_qasm
    prog = parse_qasm(C.teleport_qasm())
    with BatchRunner(workers=1 << 16, device=0) as br:
        keys, _, status, _, ms = br.run_shots(prog.n_qubits, prog.ops, 50000, seed=1)
        print("shots kernel_ms", ms, "failed", int(status.sum()))
elif mode in ("grid160", "qft32", "grover"):
    circ = C.random_circuit(18, 160, 1) if mode == "grid160" else (C.qft(32) if mode == "qft32" else C.grover(20, 20))
    with Engine(device=0, node_capacity=1 << 20) as eng:  # bench.py's engine configuration
        eng.load(circ)
        eng.apply()
        st = eng.stats()
        print(mode, "apply_ms", st["apply_ms"], "items", st["items"], "gc", st["gc_runs"], "grid", st["grid_launches"], "thin", st["thin_gates"])

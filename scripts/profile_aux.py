"""One pass over the kernels outside the gate path, for an ncu launch list: per-shot re-simulation (k_batch_shots),
subset probabilities (k_probs_*), marginal, sampling, the sparse-exchange kernels."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
from qiskit_jku_provider_b200 import circuits as C
from qiskit_jku_provider_b200.batch import BatchRunner
from qiskit_jku_provider_b200.engine import Engine, OP_DTYPE, OP_PROBABILITIES, circuit_ops
from qiskit_jku_provider_b200.qasm import parse_qasm

prog = parse_qasm(C.teleport_qasm())
with BatchRunner(workers=1 << 16, device=0) as br:
    br.run_shots(prog.n_qubits, prog.ops, 20000, seed=1)
with Engine(device=0, node_capacity=1 << 16) as eng:
    circ = C.random_circuit(12, 60, 4)
    base = circuit_ops(circ)
    ops = np.zeros(len(base) + 2, dtype=OP_DTYPE)
    ops[:len(base)] = base
    ops[-2]["kind"], ops[-2]["ctrl_pos"], ops[-2]["arg"] = OP_PROBABILITIES, 0xFFF, 0
    ops[-1]["kind"], ops[-1]["ctrl_pos"], ops[-1]["arg"] = OP_PROBABILITIES, 0x3C5, 1
    eng.load_ops(circ.n, ops)
    eng.apply()
    eng.qubit_probability(3)
    eng.load(C.qft(26))
    eng.run(1000000, 3)
    n = eng.distinct_outcomes()
    k = torch.empty(n, dtype=torch.int64, device="cuda")
    c = torch.empty(n, dtype=torch.int64, device="cuda")
    eng.counts_partition(4, k.data_ptr(), c.data_ptr(), n)
    torch.cuda.synchronize()
    eng.counts_merge(k.data_ptr(), c.data_ptr(), n, expect=n, reset=True)
print("ok")

// teleport |1> from q[0] to q[2]: mid-circuit measurements, conditioned corrections (per-shot re-simulation)
OPENQASM 2.0;
include "qelib1.inc";
qreg q[3];
creg c0[1];
creg c1[1];
creg out[1];
x q[0];
h q[1];
cx q[1], q[2];
cx q[0], q[1];
h q[0];
measure q[0] -> c0[0];
measure q[1] -> c1[0];
if (c0 == 1) z q[2];
if (c1 == 1) x q[2];
measure q[2] -> out[0];

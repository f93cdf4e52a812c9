// marginal probabilities of a subset of the qubits, listed out of order
OPENQASM 2.0;
include "qelib1.inc";
qreg q[5];
ry(0.7) q[0];
ry(1.9) q[1];
cx q[0], q[2];
cx q[1], q[3];
h q[4];
crz(1.2) q[4], q[0];
cu3(0.8, 0.1, 0.5) q[2], q[4];
snapshot(0) q[3], q[0];
snapshot(1) q[4];
snapshot(2) q[2], q[4], q[1];
snapshot(7) q[0], q[1], q[2], q[3], q[4];

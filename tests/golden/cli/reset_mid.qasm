// the reference's reset projects on |0> (Simulator::ResetQubit); a gate follows, so the counts see it
OPENQASM 2.0;
include "qelib1.inc";
qreg q[3];
creg c[3];
ry(1.1) q[0];
cx q[0], q[1];
ry(0.4) q[2];
cx q[2], q[0];
reset q[0];
snapshot(0) q[0],q[1],q[2];
x q[1];
reset q[1];
h q[2];
snapshot(1) q[0],q[1],q[2];
measure q -> c;

// every gate of qelib1 once, user gates with nested parameters and all expression operators
OPENQASM 2.0;
include "qelib1.inc";
gate mix(alpha, beta) p, r {
  u3(alpha/2, -beta, alpha*beta) p;
  cx p, r;
  u1(sin(alpha)^2 + cos(beta)) r;
  rzz(sqrt(alpha*alpha) - ln(exp(beta))) p, r;
}
gate wrap(gamma) s, t, u { mix(gamma, 2*gamma - pi/8) s, t; barrier s, t; mix(-gamma, tan(gamma/3)) u, s; ccx s, t, u; }
opaque blackbox(theta) v;
qreg q[4];
creg c[4];
u3(0.3, 1.1, -0.7) q[0];
u2(0.25, pi/3) q[1];
u1(3*pi/8) q[2];
id q[3];
u0(1) q[3];
x q[3];
y q[2];
z q[1];
h q[0];
s q[1];
sdg q[2];
t q[3];
tdg q[0];
rx(0.41) q[1];
ry(1.3) q[2];
rz(2.2) q[3];
cx q[0], q[1];
cz q[1], q[2];
cy q[2], q[3];
swap q[0], q[3];
ch q[3], q[1];
ccx q[0], q[1], q[2];
cswap q[2], q[0], q[3];
crz(0.77) q[1], q[3];
cu1(pi/5) q[2], q[0];
cu3(0.6, 0.2, 1.9) q[3], q[2];
rzz(1.05) q[0], q[2];
U(1.5e-1, .5, 2.) q[1];
CX q[3], q[0];
mix(0.9, 1.7) q[2], q[1];
wrap(0.35) q[0], q[2], q[3];
blackbox(0.1) q[1];
snapshot(0) q[0],q[1],q[2],q[3];
snapshot(1) q[3],q[1];
measure q -> c;

// whole-register arguments: h on every qubit of a, pairwise cx a -> b, single control onto a register
OPENQASM 2.0;
include "qelib1.inc";
qreg a[3];
qreg b[3];
qreg anc[1];
creg ma[3];
creg mb[3];
h a;
cx a, b;
x anc[0];
cx anc[0], b;
barrier a;
snapshot(3) a[0],a[1],a[2],b[0],b[1],b[2],anc[0];
measure a -> ma;
measure b -> mb;

// basis-state permutation: exactly one outcome (cf. test_multi_registers_convention of the reference)
OPENQASM 2.0;
include "qelib1.inc";
qreg lo[2];
qreg hi[2];
creg m[4];
x lo[1];
x hi[0];
swap lo[0], hi[1];
ccx lo[1], hi[0], lo[0];
measure lo[0] -> m[0];
measure lo[1] -> m[1];
measure hi[0] -> m[2];
measure hi[1] -> m[3];

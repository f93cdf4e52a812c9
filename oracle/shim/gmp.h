/* oracle scaffolding (test infrastructure, NOT product code).
 * Declaration-only stand-in for <gmp.h>: the reference (src/QMDDcomplex.h:35)
 * includes it but only needs the limb type that __mpfr_struct points at.
 * The real libgmp.so.10 runtime in the image is linked; no GMP call is made
 * directly by the reference sources. */
#ifndef JKQ_ORACLE_SHIM_GMP_H
#define JKQ_ORACLE_SHIM_GMP_H
typedef unsigned long int mp_limb_t;
#endif

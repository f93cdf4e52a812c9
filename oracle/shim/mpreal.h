/* oracle scaffolding (test infrastructure, NOT product code).
 * Minimal stand-in for the un-vendored third-party header "MPFR C++" (mpreal.h,
 * P. Holoborodko) that the reference includes (src/QMDDcomplex.h:34) but does not
 * ship.  Written from the documented behaviour of that class, only the members the
 * reference uses: value type over one mpfr_t, constructed at the current default
 * precision, assignment adopts the source precision, binary operators work at the
 * larger operand precision, round-to-nearest everywhere. */
#ifndef JKQ_ORACLE_SHIM_MPREAL_H
#define JKQ_ORACLE_SHIM_MPREAL_H
#include "mpfr.h"
#include <ostream>

namespace mpfr {

class mpreal {
  mpfr_t mp;
public:
  static void set_default_prec(mp_prec_t p) { mpfr_set_default_prec(p); }
  static mp_prec_t get_default_prec() { return mpfr_get_default_prec(); }

  mpreal() { mpfr_init2(mp, mpfr_get_default_prec()); mpfr_set_si(mp, 0, MPFR_RNDN); }
  mpreal(const mpreal &o) { mpfr_init2(mp, mpfr_get_prec(o.mp)); mpfr_set(mp, o.mp, MPFR_RNDN); }
  mpreal(double v) { mpfr_init2(mp, mpfr_get_default_prec()); mpfr_set_d(mp, v, MPFR_RNDN); }
  mpreal(int v) { mpfr_init2(mp, mpfr_get_default_prec()); mpfr_set_si(mp, v, MPFR_RNDN); }
  mpreal(long v) { mpfr_init2(mp, mpfr_get_default_prec()); mpfr_set_si(mp, v, MPFR_RNDN); }
  mpreal(unsigned int v) { mpfr_init2(mp, mpfr_get_default_prec()); mpfr_set_si(mp, (long)v, MPFR_RNDN); }
  mpreal(unsigned long v) { mpfr_init2(mp, mpfr_get_default_prec()); mpfr_set_d(mp, (double)v, MPFR_RNDN); }
  mpreal(long long v) { mpfr_init2(mp, mpfr_get_default_prec()); mpfr_set_si(mp, (long)v, MPFR_RNDN); }
  mpreal(unsigned long long v) { mpfr_init2(mp, mpfr_get_default_prec()); mpfr_set_d(mp, (double)v, MPFR_RNDN); }
  mpreal(mp_prec_t prec, int /*tag: zero at given precision*/, int) { mpfr_init2(mp, prec); mpfr_set_si(mp, 0, MPFR_RNDN); }
  ~mpreal() { mpfr_clear(mp); }

  mpreal &operator=(const mpreal &o) {
    if (this != &o) {
      if (mpfr_get_prec(mp) != mpfr_get_prec(o.mp)) { mpfr_clear(mp); mpfr_init2(mp, mpfr_get_prec(o.mp)); }
      mpfr_set(mp, o.mp, MPFR_RNDN);
    }
    return *this;
  }

  ::mpfr_ptr mpfr_ptr() { return mp; }
  ::mpfr_srcptr mpfr_ptr() const { return mp; }
  ::mpfr_srcptr mpfr_srcptr() const { return mp; }
  double toDouble() const { return mpfr_get_d(mp, MPFR_RNDN); }
  mp_prec_t get_prec() const { return mpfr_get_prec(mp); }

  mpreal &operator+=(const mpreal &o);
  mpreal &operator-=(const mpreal &o);
  mpreal &operator*=(const mpreal &o);
  mpreal &operator/=(const mpreal &o);
  mpreal operator-() const { mpreal r(*this); mpfr_neg(r.mp, r.mp, MPFR_RNDN); return r; }
};

namespace detail {
inline mp_prec_t pmax(const mpreal &a, const mpreal &b) { return a.get_prec() > b.get_prec() ? a.get_prec() : b.get_prec(); }
}

#define JKQ_MPREAL_BINOP(OP, FN)                                                        \
  inline mpreal operator OP(const mpreal &a, const mpreal &b) {                        \
    mpreal r(detail::pmax(a, b), 0, 0);                                                 \
    FN(r.mpfr_ptr(), a.mpfr_srcptr(), b.mpfr_srcptr(), MPFR_RNDN);                      \
    return r;                                                                           \
  }                                                                                     \
  template <class T> inline mpreal operator OP(const mpreal &a, const T &b) { return a OP mpreal(b); } \
  template <class T> inline mpreal operator OP(const T &a, const mpreal &b) { return mpreal(a) OP b; }
JKQ_MPREAL_BINOP(+, mpfr_add)
JKQ_MPREAL_BINOP(-, mpfr_sub)
JKQ_MPREAL_BINOP(*, mpfr_mul)
JKQ_MPREAL_BINOP(/, mpfr_div)
#undef JKQ_MPREAL_BINOP

inline mpreal &mpreal::operator+=(const mpreal &o) { *this = *this + o; return *this; }
inline mpreal &mpreal::operator-=(const mpreal &o) { *this = *this - o; return *this; }
inline mpreal &mpreal::operator*=(const mpreal &o) { *this = *this * o; return *this; }
inline mpreal &mpreal::operator/=(const mpreal &o) { *this = *this / o; return *this; }

#define JKQ_MPREAL_CMP(OP)                                                              \
  inline bool operator OP(const mpreal &a, const mpreal &b) { return mpfr_cmp(a.mpfr_srcptr(), b.mpfr_srcptr()) OP 0; } \
  template <class T> inline bool operator OP(const mpreal &a, const T &b) { return a OP mpreal(b); } \
  template <class T> inline bool operator OP(const T &a, const mpreal &b) { return mpreal(a) OP b; }
JKQ_MPREAL_CMP(<)
JKQ_MPREAL_CMP(>)
JKQ_MPREAL_CMP(<=)
JKQ_MPREAL_CMP(>=)
JKQ_MPREAL_CMP(==)
JKQ_MPREAL_CMP(!=)
#undef JKQ_MPREAL_CMP

#define JKQ_MPREAL_FN1(NAME, FN)                                                        \
  inline mpreal NAME(const mpreal &a) {                                                 \
    mpreal r(a.get_prec(), 0, 0);                                                       \
    FN(r.mpfr_ptr(), a.mpfr_srcptr(), MPFR_RNDN);                                       \
    return r;                                                                           \
  }
JKQ_MPREAL_FN1(sqrt, mpfr_sqrt)
JKQ_MPREAL_FN1(cos, mpfr_cos)
JKQ_MPREAL_FN1(sin, mpfr_sin)
JKQ_MPREAL_FN1(tan, mpfr_tan)
JKQ_MPREAL_FN1(exp, mpfr_exp)
JKQ_MPREAL_FN1(log, mpfr_log)
JKQ_MPREAL_FN1(acos, mpfr_acos)
JKQ_MPREAL_FN1(abs, mpfr_abs)
JKQ_MPREAL_FN1(fabs, mpfr_abs)
#undef JKQ_MPREAL_FN1

inline mpreal pow(const mpreal &a, const mpreal &b) {
  mpreal r(detail::pmax(a, b), 0, 0);
  mpfr_pow(r.mpfr_ptr(), a.mpfr_srcptr(), b.mpfr_srcptr(), MPFR_RNDN);
  return r;
}

inline mpreal const_pi() {
  mpreal r;
  mpfr_const_pi(r.mpfr_ptr(), MPFR_RNDN);
  return r;
}

inline std::ostream &operator<<(std::ostream &os, const mpreal &v) { return os << v.toDouble(); }

}  // namespace mpfr
#endif

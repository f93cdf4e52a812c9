// jkq_qasm.cpp -- OpenQASM 2.0 front end of the engine: text -> array of jkq_op (host code, no CUDA).
//
// Replaces the executing parser of the reference (src/QASMsimulator.cpp:895-1060 on top of src/QASMscanner.cpp):
// the reference applies every gate to the decision diagram while it parses; here parsing produces the op list
// that jkq_load_circuit takes, so a program is parsed once and can be run any number of times (per-shot
// re-simulation, QASMsimulator.cpp:501-522, parses the text again for every shot).
// Semantics kept from the reference:
//   * U(theta,phi,lambda) in the SU(2) convention of QASMsimulator.cpp:295-298; CX = exact X with one control
//   * user gates are flattened to U / CX when they are declared (:682-757); opaque gates have an empty body
//   * register arguments are broadcast per primitive of the gate body, not per call (:394-452)
//   * `measure` collapses immediately (Simulator::MeasureOne), `reset` is Simulator::ResetQubit, `if (c == k) qop`
//     guards the whole qop, `snapshot(k) q...` records probabilities / the state vector, `show_probabilities`
//   * qubit q of the first qreg is qubit 0; registers are laid out in declaration order (:909-910)
// Deliberate differences: real literals such as 1e-05 are read (QASMscanner.cpp:249-283 cannot); a syntax error
// is reported with line and column and stops the parse (the reference prints to stderr and carries on);
// `CX reg, q[0]` loops over the control register (:329-337 applies the first control only).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/jkq.h"

namespace {

// the standard header: used for `include "qelib1.inc";` when no such file is found next to the program
const char *const QELIB1 =
    "gate u3(theta,phi,lambda) q { U(theta,phi,lambda) q; }\n"
    "gate u2(phi,lambda) q { U(pi/2,phi,lambda) q; }\n"
    "gate u1(lambda) q { U(0,0,lambda) q; }\n"
    "gate cx c,t { CX c,t; }\n"
    "gate id a { U(0,0,0) a; }\n"
    "gate u0(gamma) q { U(0,0,0) q; }\n"
    "gate x a { u3(pi,0,pi) a; }\n"
    "gate y a { u3(pi,pi/2,pi/2) a; }\n"
    "gate z a { u1(pi) a; }\n"
    "gate h a { u2(0,pi) a; }\n"
    "gate s a { u1(pi/2) a; }\n"
    "gate sdg a { u1(-pi/2) a; }\n"
    "gate t a { u1(pi/4) a; }\n"
    "gate tdg a { u1(-pi/4) a; }\n"
    "gate rx(theta) a { u3(theta,-pi/2,pi/2) a; }\n"
    "gate ry(theta) a { u3(theta,0,0) a; }\n"
    "gate rz(phi) a { u1(phi) a; }\n"
    "gate cz a,b { h b; cx a,b; h b; }\n"
    "gate cy a,b { sdg b; cx a,b; s b; }\n"
    "gate swap a,b { cx a,b; cx b,a; cx a,b; }\n"
    "gate ch a,b { h b; sdg b; cx a,b; h b; t b; cx a,b; t b; h b; s b; x b; s a; }\n"
    "gate ccx a,b,c { h c; cx b,c; tdg c; cx a,c; t c; cx b,c; tdg c; cx a,c; t b; t c; h c; cx a,b; t a; tdg b; cx a,b; }\n"
    "gate cswap a,b,c { cx c,b; ccx a,b,c; cx c,b; }\n"
    "gate crz(lambda) a,b { u1(lambda/2) b; cx a,b; u1(-lambda/2) b; cx a,b; }\n"
    "gate cu1(lambda) a,b { u1(lambda/2) a; cx a,b; u1(-lambda/2) b; cx a,b; u1(lambda/2) b; }\n"
    "gate cu3(theta,phi,lambda) c,t { u1((lambda-phi)/2) t; cx c,t; u3(-theta/2,0,-(phi+lambda)/2) t; cx c,t; u3(theta/2,phi,0) t; }\n"
    "gate rzz(theta) a,b { cx a,b; u1(theta) b; cx a,b; }\n";

enum Tok {
  T_EOF, T_ID, T_INT, T_REAL, T_STR, T_SEMI, T_COMMA, T_LPAR, T_RPAR, T_LBRACK, T_RBRACK, T_LBRACE, T_RBRACE,
  T_PLUS, T_MINUS, T_TIMES, T_DIV, T_POW, T_GT, T_EQ, T_BAD
};

struct Token {
  Tok kind = T_EOF;
  std::string str;
  double num = 0;
  int line = 1, col = 1;
};

struct Source {
  std::string text, name;
  size_t pos = 0;
  int line = 1, col = 1;
};

struct ParseError {
  std::string msg;
};

class Lexer {
 public:
  void push(const std::string &text, const std::string &name) {
    Source s;
    s.text = text; s.name = name;
    stack_.push_back(s);
  }
  Token next() {
    for (;;) {
      if (stack_.empty()) return Token();
      Source &s = stack_.back();
      skip_blank(s);
      if (s.pos >= s.text.size()) {
        if (stack_.size() == 1) { Token t; t.line = s.line; t.col = s.col; return t; }
        stack_.pop_back();  // end of an included file: continue behind the include statement
        continue;
      }
      return scan(s);
    }
  }
  std::string where(const Token &t) const {
    char buf[96];
    snprintf(buf, sizeof(buf), "line %d, column %d", t.line, t.col);
    std::string w = buf;
    if (stack_.size() > 1) w += " of " + stack_.back().name;
    return w;
  }

 private:
  static void advance(Source &s) {
    if (s.text[s.pos] == '\n') { s.line++; s.col = 1; }
    else s.col++;
    s.pos++;
  }
  static void skip_blank(Source &s) {
    for (;;) {
      while (s.pos < s.text.size() && isspace((unsigned char)s.text[s.pos])) advance(s);
      if (s.pos + 1 < s.text.size() && s.text[s.pos] == '/' && s.text[s.pos + 1] == '/') {
        while (s.pos < s.text.size() && s.text[s.pos] != '\n') advance(s);
        continue;
      }
      return;
    }
  }
  static Token scan(Source &s) {
    Token t;
    t.line = s.line; t.col = s.col;
    const char ch = s.text[s.pos];
    auto peek = [&](size_t k) { return s.pos + k < s.text.size() ? s.text[s.pos + k] : '\0'; };
    if (isalpha((unsigned char)ch) || ch == '_') {
      while (s.pos < s.text.size() && (isalnum((unsigned char)s.text[s.pos]) || s.text[s.pos] == '_')) { t.str += s.text[s.pos]; advance(s); }
      t.kind = T_ID;
      return t;
    }
    if (isdigit((unsigned char)ch) || (ch == '.' && isdigit((unsigned char)peek(1)))) {
      bool real = false;
      while (isdigit((unsigned char)peek(0))) { t.str += peek(0); advance(s); }
      if (peek(0) == '.') {
        real = true;
        t.str += '.'; advance(s);
        while (isdigit((unsigned char)peek(0))) { t.str += peek(0); advance(s); }
      }
      if ((peek(0) == 'e' || peek(0) == 'E') &&
          (isdigit((unsigned char)peek(1)) || ((peek(1) == '+' || peek(1) == '-') && isdigit((unsigned char)peek(2))))) {
        real = true;
        t.str += peek(0); advance(s);
        if (peek(0) == '+' || peek(0) == '-') { t.str += peek(0); advance(s); }
        while (isdigit((unsigned char)peek(0))) { t.str += peek(0); advance(s); }
      }
      t.kind = real ? T_REAL : T_INT;
      t.num = strtod(t.str.c_str(), nullptr);
      return t;
    }
    if (ch == '"') {
      advance(s);
      while (s.pos < s.text.size() && s.text[s.pos] != '"') { t.str += s.text[s.pos]; advance(s); }
      if (s.pos < s.text.size()) advance(s);
      t.kind = T_STR;
      return t;
    }
    advance(s);
    switch (ch) {
      case ';': t.kind = T_SEMI; break;
      case ',': t.kind = T_COMMA; break;
      case '(': t.kind = T_LPAR; break;
      case ')': t.kind = T_RPAR; break;
      case '[': t.kind = T_LBRACK; break;
      case ']': t.kind = T_RBRACK; break;
      case '{': t.kind = T_LBRACE; break;
      case '}': t.kind = T_RBRACE; break;
      case '+': t.kind = T_PLUS; break;
      case '-': t.kind = T_MINUS; break;
      case '*': t.kind = T_TIMES; break;
      case '/': t.kind = T_DIV; break;
      case '^': t.kind = T_POW; break;
      case '>': t.kind = T_GT; break;
      case '=':
        if (s.pos < s.text.size() && s.text[s.pos] == '=') { advance(s); t.kind = T_EQ; }
        else { t.kind = T_BAD; t.str = "="; }
        break;
      default: t.kind = T_BAD; t.str = std::string(1, ch);
    }
    return t;
  }
  std::vector<Source> stack_;
};

// ---- parameter expressions -----------------------------------------------------------------------------------
enum ExOp { X_NUM, X_PAR, X_NEG, X_ADD, X_SUB, X_MUL, X_DIV, X_POW, X_SIN, X_COS, X_TAN, X_EXP, X_LN, X_SQRT };
struct Ex {
  ExOp op;
  double num;
  int a, b;  // operand nodes (X_PAR: a = parameter index)
};
typedef std::vector<Ex> ExPool;

double ex_eval(const ExPool &pool, int i, const std::vector<double> &params) {
  const Ex &e = pool[i];
  switch (e.op) {
    case X_NUM: return e.num;
    case X_PAR: return params[e.a];
    case X_NEG: return -ex_eval(pool, e.a, params);
    case X_ADD: return ex_eval(pool, e.a, params) + ex_eval(pool, e.b, params);
    case X_SUB: return ex_eval(pool, e.a, params) - ex_eval(pool, e.b, params);
    case X_MUL: return ex_eval(pool, e.a, params) * ex_eval(pool, e.b, params);
    case X_DIV: return ex_eval(pool, e.a, params) / ex_eval(pool, e.b, params);
    case X_POW: return pow(ex_eval(pool, e.a, params), ex_eval(pool, e.b, params));
    case X_SIN: return sin(ex_eval(pool, e.a, params));
    case X_COS: return cos(ex_eval(pool, e.a, params));
    case X_TAN: return tan(ex_eval(pool, e.a, params));
    case X_EXP: return exp(ex_eval(pool, e.a, params));
    case X_LN: return log(ex_eval(pool, e.a, params));
    case X_SQRT: return sqrt(ex_eval(pool, e.a, params));
  }
  return 0;
}
int ex_new(ExPool &pool, ExOp op, double num, int a, int b) {
  Ex e;
  e.op = op; e.num = num; e.a = a; e.b = b;
  pool.push_back(e);
  return (int)pool.size() - 1;
}
// copy expression `i` of `from` into `to`, replacing parameter k by the node subst[k] of `to`
int ex_copy(const ExPool &from, int i, ExPool &to, const std::vector<int> &subst) {
  const Ex &e = from[i];
  if (e.op == X_NUM) return ex_new(to, X_NUM, e.num, -1, -1);
  if (e.op == X_PAR) return subst[e.a];
  const int a = ex_copy(from, e.a, to, subst);
  const int b = (e.op >= X_ADD && e.op <= X_POW) ? ex_copy(from, e.b, to, subst) : -1;
  return ex_new(to, e.op, 0, a, b);
}

struct BodyOp {
  bool cx;
  int e[3];  // U: theta, phi, lambda (nodes of the gate's pool)
  int q[2];  // formal qubit arguments: U target | CX control, target
};
struct GateDef {
  std::vector<std::string> params, args;
  std::vector<BodyOp> body;  // flattened to U / CX
  ExPool pool;
};

struct Reg {
  std::string name;
  uint32_t offset, size;
};
struct SnapshotInfo {
  uint32_t slot;
  std::vector<uint32_t> qubits;  // in argument order
};

}  // namespace

struct jkq_program {
  std::string error;
  std::vector<jkq_op> ops;
  std::vector<Reg> qregs, cregs;
  std::vector<SnapshotInfo> snapshots;
  uint32_t n_qubits = 0, n_clbits = 0;
  int64_t first_measure = -1;
  int32_t gate_after_measure = 0;
};

namespace {

class Parser {
 public:
  Parser(jkq_program *p, const std::string &include_dir, uint32_t flags) : prog_(p), include_dir_(include_dir), flags_(flags) {}

  void run(const std::string &text) {
    lex_.push(text, "<input>");
    scan();
    expect_id("OPENQASM");
    if (la_.kind != T_REAL && la_.kind != T_INT) fail("expected the version number after OPENQASM");
    scan();
    expect(T_SEMI, "';'");
    while (la_.kind != T_EOF) statement();
    if (prog_->n_qubits == 0) throw ParseError{"the program declares no qreg"};
  }

 private:
  // ---- token plumbing
  void scan() { t_ = la_; la_ = lex_.next(); }
  [[noreturn]] void fail(const std::string &what) { throw ParseError{what + " (" + lex_.where(la_) + ")"}; }
  bool is_id(const char *s) const { return la_.kind == T_ID && la_.str == s; }
  void expect(Tok k, const char *name) {
    if (la_.kind != k) fail(std::string("expected ") + name + " but found '" + show(la_) + "'");
    scan();
  }
  void expect_id(const char *s) {
    if (!is_id(s)) fail(std::string("expected '") + s + "' but found '" + show(la_) + "'");
    scan();
  }
  static std::string show(const Token &t) {
    if (t.kind == T_EOF) return "end of input";
    if (!t.str.empty()) return t.str;
    static const char *names[] = {"eof", "id", "int", "real", "string", ";", ",", "(", ")", "[", "]", "{", "}", "+", "-", "*", "/", "^", ">", "==", "?"};
    return names[t.kind];
  }
  std::string ident() {
    if (la_.kind != T_ID) fail("expected an identifier but found '" + show(la_) + "'");
    scan();
    return t_.str;
  }
  uint32_t nninteger() {
    if (la_.kind != T_INT) fail("expected a non-negative integer but found '" + show(la_) + "'");
    scan();
    return (uint32_t)t_.num;
  }

  // ---- expressions: exp := ['-'] term {(+|-) term}; term := factor {(*|/) factor}; factor := atom {^ atom}
  int atom(ExPool &pool, const std::vector<std::string> *names) {
    if (la_.kind == T_REAL || la_.kind == T_INT) { scan(); return ex_new(pool, X_NUM, t_.num, -1, -1); }
    if (la_.kind == T_MINUS) { scan(); return ex_new(pool, X_NEG, 0, atom(pool, names), -1); }
    if (la_.kind == T_LPAR) {
      scan();
      const int x = expr(pool, names);
      expect(T_RPAR, "')'");
      return x;
    }
    if (la_.kind == T_ID) {
      static const struct { const char *name; ExOp op; } fns[] = {{"sin", X_SIN}, {"cos", X_COS}, {"tan", X_TAN}, {"exp", X_EXP}, {"ln", X_LN}, {"sqrt", X_SQRT}};
      for (const auto &f : fns) {
        if (la_.str != f.name) continue;
        scan();
        expect(T_LPAR, "'('");
        const int x = expr(pool, names);
        expect(T_RPAR, "')'");
        return ex_new(pool, f.op, 0, x, -1);
      }
      if (la_.str == "pi") { scan(); return ex_new(pool, X_NUM, M_PI, -1, -1); }
      if (names) {
        for (size_t k = 0; k < names->size(); k++)
          if ((*names)[k] == la_.str) { scan(); return ex_new(pool, X_PAR, 0, (int)k, -1); }
      }
      fail("unknown identifier '" + la_.str + "' in an expression");
    }
    fail("invalid expression at '" + show(la_) + "'");
  }
  int factor(ExPool &pool, const std::vector<std::string> *names) {
    int x = atom(pool, names);
    while (la_.kind == T_POW) { scan(); x = ex_new(pool, X_POW, 0, x, atom(pool, names)); }
    return x;
  }
  int term(ExPool &pool, const std::vector<std::string> *names) {
    int x = factor(pool, names);
    while (la_.kind == T_TIMES || la_.kind == T_DIV) {
      const ExOp op = la_.kind == T_TIMES ? X_MUL : X_DIV;
      scan();
      x = ex_new(pool, op, 0, x, factor(pool, names));
    }
    return x;
  }
  int expr(ExPool &pool, const std::vector<std::string> *names) {
    int x;
    if (la_.kind == T_MINUS) { scan(); x = ex_new(pool, X_NEG, 0, term(pool, names), -1); }
    else x = term(pool, names);
    while (la_.kind == T_PLUS || la_.kind == T_MINUS) {
      const ExOp op = la_.kind == T_PLUS ? X_ADD : X_SUB;
      scan();
      x = ex_new(pool, op, 0, x, term(pool, names));
    }
    return x;
  }
  double const_expr() {
    ExPool pool;
    const int x = expr(pool, nullptr);
    return ex_eval(pool, x, std::vector<double>());
  }

  // ---- arguments
  struct QArg { uint32_t first, size; };
  QArg qarg() {
    const std::string name = ident();
    const Reg *r = nullptr;
    for (const Reg &q : prog_->qregs) if (q.name == name) r = &q;
    if (!r) fail("argument is not a qreg: " + name);
    if (la_.kind == T_LBRACK) {
      scan();
      const uint32_t i = nninteger();
      expect(T_RBRACK, "']'");
      if (i >= r->size) fail("index of qreg " + name + " is out of bounds");
      return QArg{r->offset + i, 1};
    }
    return QArg{r->offset, r->size};
  }
  std::vector<QArg> qarg_list() {
    std::vector<QArg> v;
    v.push_back(qarg());
    while (la_.kind == T_COMMA) { scan(); v.push_back(qarg()); }
    return v;
  }
  std::vector<std::string> id_list() {
    std::vector<std::string> v;
    v.push_back(ident());
    while (la_.kind == T_COMMA) { scan(); v.push_back(ident()); }
    return v;
  }

  // ---- op emission
  void emit(const jkq_op &o) {
    if (o.kind == JKQ_OP_GATE && prog_->first_measure >= 0) prog_->gate_after_measure = 1;
    if (o.kind == JKQ_OP_MEASURE_ONE && prog_->first_measure < 0) prog_->first_measure = (int64_t)prog_->ops.size();
    prog_->ops.push_back(o);
  }
  static jkq_op blank(uint32_t kind) {
    jkq_op o;
    memset(&o, 0, sizeof(o));
    o.kind = kind;
    return o;
  }
  void emit_u(double theta, double phi, double lambda, uint32_t qubit) {
    // QASMsimulator.cpp:295-298
    jkq_op o = blank(JKQ_OP_GATE);
    o.target = qubit;
    const double c = cos(theta / 2), s = sin(theta / 2), p = (phi + lambda) / 2, m = (phi - lambda) / 2;
    o.m[0] = cos(-p) * c;  o.m[1] = sin(-p) * c;
    o.m[2] = -cos(-m) * s; o.m[3] = -sin(-m) * s;
    o.m[4] = cos(m) * s;   o.m[5] = sin(m) * s;
    o.m[6] = cos(p) * c;   o.m[7] = sin(p) * c;
    emit(o);
  }
  void emit_cx(uint32_t control, uint32_t target) {
    if (control == target) fail("CX with identical control and target");
    jkq_op o = blank(JKQ_OP_GATE);
    o.target = target;
    o.ctrl_pos = 1ull << control;
    o.m[2] = 1.0; o.m[4] = 1.0;
    emit(o);
  }
  void broadcast_u(double th, double ph, double la, QArg t) {
    for (uint32_t i = 0; i < t.size; i++) emit_u(th, ph, la, t.first + i);
  }
  void broadcast_cx(QArg c, QArg t) {
    if (c.size == t.size) { for (uint32_t i = 0; i < t.size; i++) emit_cx(c.first + i, t.first + i); }
    else if (c.size == 1) { for (uint32_t i = 0; i < t.size; i++) emit_cx(c.first, t.first + i); }
    else if (t.size == 1) { for (uint32_t i = 0; i < c.size; i++) emit_cx(c.first + i, t.first); }
    else fail("register sizes do not match for CX");
  }

  // ---- quantum operations (QASMqop): returns through emit(); `guard` wraps what is emitted into an IF
  void qop() {
    if (is_id("U")) {
      scan();
      expect(T_LPAR, "'('");
      const double th = const_expr();
      expect(T_COMMA, "','");
      const double ph = const_expr();
      expect(T_COMMA, "','");
      const double la = const_expr();
      expect(T_RPAR, "')'");
      const QArg t = qarg();
      expect(T_SEMI, "';'");
      broadcast_u(th, ph, la, t);
    } else if (is_id("CX")) {
      scan();
      const QArg c = qarg();
      expect(T_COMMA, "','");
      const QArg t = qarg();
      expect(T_SEMI, "';'");
      broadcast_cx(c, t);
    } else if (is_id("measure")) {
      scan();
      const QArg q = qarg();
      expect(T_MINUS, "'->'");
      expect(T_GT, "'->'");
      const std::string cname = ident();
      const Reg *r = nullptr;
      for (const Reg &c : prog_->cregs) if (c.name == cname) r = &c;
      if (!r) fail("argument is not a creg: " + cname);
      uint32_t cfirst = r->offset, csize = r->size;
      if (la_.kind == T_LBRACK) {
        scan();
        const uint32_t i = nninteger();
        expect(T_RBRACK, "']'");
        if (i >= r->size) fail("index of creg " + cname + " is out of bounds");
        cfirst += i; csize = 1;
      }
      expect(T_SEMI, "';'");
      if (q.size != csize) fail("mismatch of qreg and creg size in measurement");
      for (uint32_t i = 0; i < csize; i++) {
        if (cfirst + i >= 64) fail("measurements into more than 64 classical bits are not supported");
        jkq_op o = blank(JKQ_OP_MEASURE_ONE);
        o.target = q.first + i;
        o.arg = cfirst + i;
        emit(o);
      }
    } else if (is_id("reset")) {
      scan();
      const QArg q = qarg();
      expect(T_SEMI, "';'");
      for (uint32_t i = 0; i < q.size; i++) {
        jkq_op o = blank(JKQ_OP_RESET);
        o.target = q.first + i;
        emit(o);
      }
    } else if (la_.kind == T_ID) {
      const std::string name = ident();
      auto it = gates_.find(name);
      if (it == gates_.end()) fail("undefined gate: " + name);
      const GateDef &g = it->second;
      std::vector<double> params;
      if (la_.kind == T_LPAR) {
        scan();
        if (la_.kind != T_RPAR) {
          params.push_back(const_expr());
          while (la_.kind == T_COMMA) { scan(); params.push_back(const_expr()); }
        }
        expect(T_RPAR, "')'");
      }
      const std::vector<QArg> args = qarg_list();
      expect(T_SEMI, "';'");
      if (params.size() != g.params.size()) fail("gate " + name + " takes a different number of parameters");
      if (args.size() != g.args.size()) fail("gate " + name + " takes a different number of qubit arguments");
      uint32_t size = 1;
      for (const QArg &a : args) {
        if (a.size > 1 && size != 1 && a.size != size) fail("register sizes do not match in the call of " + name);
        if (a.size > 1) size = a.size;
      }
      for (const BodyOp &b : g.body) {
        if (b.cx) broadcast_cx(args[b.q[0]], args[b.q[1]]);
        else broadcast_u(ex_eval(g.pool, b.e[0], params), ex_eval(g.pool, b.e[1], params), ex_eval(g.pool, b.e[2], params), args[b.q[0]]);
      }
    } else {
      fail("expected a quantum operation but found '" + show(la_) + "'");
    }
  }

  // ---- gate declarations
  static int formal(const std::vector<std::string> &names, const std::string &s) {
    for (size_t i = 0; i < names.size(); i++) if (names[i] == s) return (int)i;
    return -1;
  }
  void gate_decl(bool opaque) {
    scan();  // 'gate' / 'opaque'
    const std::string name = ident();
    GateDef g;
    if (la_.kind == T_LPAR) {
      scan();
      if (la_.kind != T_RPAR) g.params = id_list();
      expect(T_RPAR, "')'");
    }
    g.args = id_list();
    if (opaque) {  // empty body, QASMsimulator.cpp:662-680
      expect(T_SEMI, "';'");
      gates_[name] = g;
      return;
    }
    expect(T_LBRACE, "'{'");
    auto arg_of = [&](const std::string &s) {
      const int k = formal(g.args, s);
      if (k < 0) fail("'" + s + "' is not an argument of gate " + name);
      return k;
    };
    while (la_.kind != T_RBRACE) {
      if (la_.kind == T_EOF) fail("unterminated body of gate " + name);
      if (is_id("U")) {
        scan();
        BodyOp b;
        b.cx = false; b.q[1] = -1;
        expect(T_LPAR, "'('");
        b.e[0] = expr(g.pool, &g.params);
        expect(T_COMMA, "','");
        b.e[1] = expr(g.pool, &g.params);
        expect(T_COMMA, "','");
        b.e[2] = expr(g.pool, &g.params);
        expect(T_RPAR, "')'");
        b.q[0] = arg_of(ident());
        expect(T_SEMI, "';'");
        g.body.push_back(b);
      } else if (is_id("CX")) {
        scan();
        BodyOp b;
        b.cx = true; b.e[0] = b.e[1] = b.e[2] = -1;
        b.q[0] = arg_of(ident());
        expect(T_COMMA, "','");
        b.q[1] = arg_of(ident());
        expect(T_SEMI, "';'");
        g.body.push_back(b);
      } else if (is_id("barrier")) {
        scan();
        id_list();
        expect(T_SEMI, "';'");
      } else if (la_.kind == T_ID) {
        const std::string callee = ident();
        auto it = gates_.find(callee);
        if (it == gates_.end()) fail("undefined gate: " + callee);
        const GateDef &cg = it->second;
        std::vector<int> subst;
        if (la_.kind == T_LPAR) {
          scan();
          if (la_.kind != T_RPAR) {
            subst.push_back(expr(g.pool, &g.params));
            while (la_.kind == T_COMMA) { scan(); subst.push_back(expr(g.pool, &g.params)); }
          }
          expect(T_RPAR, "')'");
        }
        const std::vector<std::string> actual = id_list();
        expect(T_SEMI, "';'");
        if (subst.size() != cg.params.size()) fail("gate " + callee + " takes a different number of parameters");
        if (actual.size() != cg.args.size()) fail("gate " + callee + " takes a different number of qubit arguments");
        std::vector<int> amap;
        for (const std::string &s : actual) amap.push_back(arg_of(s));
        for (const BodyOp &cb : cg.body) {  // inline the callee's flattened body (QASMsimulator.cpp:736-757)
          BodyOp b = cb;
          b.q[0] = amap[cb.q[0]];
          if (cb.cx) b.q[1] = amap[cb.q[1]];
          else for (int k = 0; k < 3; k++) b.e[k] = ex_copy(cg.pool, cb.e[k], g.pool, subst);
          g.body.push_back(b);
        }
      } else {
        fail("error in the declaration of gate " + name + " at '" + show(la_) + "'");
      }
    }
    scan();  // '}'
    gates_[name] = g;
  }

  static std::string read_file(const std::string &path, bool *ok) {
    std::ifstream f(path.c_str(), std::ios::in | std::ios::binary);
    *ok = (bool)f;
    std::stringstream ss;
    if (f) ss << f.rdbuf();
    return ss.str();
  }

  void statement() {
    if (is_id("qreg") || is_id("creg")) {
      const bool q = la_.str == "qreg";
      scan();
      Reg r;
      r.name = ident();
      expect(T_LBRACK, "'['");
      r.size = nninteger();
      expect(T_RBRACK, "']'");
      expect(T_SEMI, "';'");
      if (r.size == 0) fail("register " + r.name + " has size 0");
      if (q) {
        r.offset = prog_->n_qubits;
        prog_->n_qubits += r.size;
        if (prog_->n_qubits > 64) fail("more than 64 qubits");
        prog_->qregs.push_back(r);
      } else {
        r.offset = prog_->n_clbits;
        prog_->n_clbits += r.size;
        prog_->cregs.push_back(r);
      }
    } else if (is_id("gate")) {
      gate_decl(false);
    } else if (is_id("opaque")) {
      gate_decl(true);
    } else if (is_id("include")) {
      scan();
      if (la_.kind != T_STR) fail("expected a file name after include");
      const std::string fname = la_.str;
      scan();
      if (la_.kind != T_SEMI) fail("expected ';' after include");
      // (the next token is read from the included text: push it before scanning on)
      bool ok = false;
      std::string text;
      if (!include_dir_.empty()) text = read_file(include_dir_ + "/" + fname, &ok);
      if (!ok) text = read_file(fname, &ok);
      if (!ok && fname == "qelib1.inc") { text = QELIB1; ok = true; }
      if (!ok) fail("failed to open file '" + fname + "'");
      lex_.push(text, fname);
      scan();
    } else if (is_id("barrier")) {
      scan();
      qarg_list();
      expect(T_SEMI, "';'");
      emit(blank(JKQ_OP_BARRIER));
    } else if (is_id("if")) {
      scan();
      expect(T_LPAR, "'('");
      const std::string cname = ident();
      const Reg *r = nullptr;
      for (const Reg &c : prog_->cregs) if (c.name == cname) r = &c;
      if (!r) fail("error in if statement: " + cname + " is not a creg");
      expect(T_EQ, "'=='");
      if (la_.kind != T_INT) fail("expected a non-negative integer in the if condition");
      scan();
      const double value = t_.num;
      expect(T_RPAR, "')'");
      if (r->offset + r->size > 64) fail("conditions on classical bits beyond 64 are not supported");
      const size_t at = prog_->ops.size();
      jkq_op o = blank(JKQ_OP_IF);
      o.ctrl_pos = (r->size == 64 ? ~0ull : ((1ull << r->size) - 1)) << r->offset;
      // a value that does not fit the register can never match: leave a bit outside the mask set
      if (value >= 18446744073709551616.0 || (r->size < 64 && (uint64_t)value >> r->size)) o.ctrl_neg = ~o.ctrl_pos;
      else o.ctrl_neg = (uint64_t)value << r->offset;
      if (o.ctrl_neg & ~o.ctrl_pos) o.ctrl_neg = ~0ull;
      prog_->ops.push_back(o);
      qop();
      const size_t guarded = prog_->ops.size() - at - 1;
      if (guarded == 0) prog_->ops.pop_back();
      else prog_->ops[at].arg = (uint32_t)guarded;
    } else if (is_id("snapshot")) {
      scan();
      expect(T_LPAR, "'('");
      SnapshotInfo s;
      s.slot = nninteger();
      expect(T_RPAR, "')'");
      const std::vector<QArg> args = qarg_list();
      expect(T_SEMI, "';'");
      uint64_t mask = 0;
      for (const QArg &a : args) {
        if (a.size != 1) fail("error in snapshot: arguments must be qubits");
        if (mask >> a.first & 1) fail("error in snapshot: a qubit occurs twice");
        mask |= 1ull << a.first;
        s.qubits.push_back(a.first);
      }
      prog_->snapshots.push_back(s);
      if (flags_ & JKQ_QASM_PROBABILITIES) {
        jkq_op o = blank(JKQ_OP_PROBABILITIES);
        o.ctrl_pos = mask;
        o.arg = s.slot;
        emit(o);
      }
      if ((flags_ & JKQ_QASM_STATEVECTOR) && s.qubits.size() == prog_->n_qubits) {
        jkq_op o = blank(JKQ_OP_SNAPSHOT);
        o.arg = s.slot;
        emit(o);
      }
    } else if (is_id("show_probabilities")) {
      scan();
      expect(T_SEMI, "';'");
      jkq_op o = blank(JKQ_OP_PROBABILITIES);
      o.ctrl_pos = prog_->n_qubits == 64 ? ~0ull : ((1ull << prog_->n_qubits) - 1);
      o.arg = JKQ_SLOT_SHOW_PROBABILITIES + n_show_++;
      emit(o);
    } else if (la_.kind == T_ID) {
      qop();
    } else {
      fail("unexpected statement: started with '" + show(la_) + "'");
    }
  }

  jkq_program *prog_;
  std::string include_dir_;
  uint32_t flags_;
  uint32_t n_show_ = 0;
  Lexer lex_;
  Token t_, la_;
  std::map<std::string, GateDef> gates_;
};

}  // namespace

extern "C" int jkq_qasm_parse(const char *text, const char *include_dir, uint32_t flags, jkq_program **out) {
  if (!text || !out) return JKQ_E_ARGUMENT;
  jkq_program *p = new jkq_program();
  *out = p;
  try {
    Parser parser(p, include_dir ? include_dir : "", flags);
    parser.run(text);
  } catch (const ParseError &e) {
    p->error = "error while parsing the QASM program: " + e.msg;
    p->ops.clear();
    return JKQ_E_ARGUMENT;
  }
  return JKQ_E_OK;
}

extern "C" const char *jkq_program_error(const jkq_program *p) { return p ? p->error.c_str() : "null program"; }

extern "C" void jkq_program_free(jkq_program *p) { delete p; }

extern "C" int jkq_program_get_info(const jkq_program *p, jkq_program_info *info) {
  if (!p || !info) return JKQ_E_ARGUMENT;
  info->n_qubits = p->n_qubits;
  info->n_clbits = p->n_clbits;
  info->n_ops = p->ops.size();
  info->ops = p->ops.empty() ? nullptr : p->ops.data();
  info->n_qregs = (uint32_t)p->qregs.size();
  info->n_cregs = (uint32_t)p->cregs.size();
  info->n_snapshots = (uint32_t)p->snapshots.size();
  info->gate_after_measure = p->gate_after_measure;
  info->first_measure = p->first_measure;
  return JKQ_E_OK;
}

extern "C" int jkq_program_get_register(const jkq_program *p, int creg, uint32_t index, char *name, size_t name_cap, uint32_t *offset,
                                        uint32_t *size) {
  if (!p) return JKQ_E_ARGUMENT;
  const std::vector<Reg> &v = creg ? p->cregs : p->qregs;
  if (index >= v.size()) return JKQ_E_ARGUMENT;
  if (name && name_cap) snprintf(name, name_cap, "%s", v[index].name.c_str());
  if (offset) *offset = v[index].offset;
  if (size) *size = v[index].size;
  return JKQ_E_OK;
}

extern "C" int jkq_program_get_snapshot(const jkq_program *p, uint32_t index, uint32_t *slot, uint32_t *qubits, size_t cap, size_t *n_qubits) {
  if (!p || index >= p->snapshots.size()) return JKQ_E_ARGUMENT;
  const SnapshotInfo &s = p->snapshots[index];
  if (slot) *slot = s.slot;
  if (n_qubits) *n_qubits = s.qubits.size();
  if (qubits) {
    if (cap < s.qubits.size()) return JKQ_E_CAPACITY;
    for (size_t i = 0; i < s.qubits.size(); i++) qubits[i] = s.qubits[i];
  }
  return JKQ_E_OK;
}

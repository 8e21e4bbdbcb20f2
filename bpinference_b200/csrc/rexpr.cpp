// rexpr.cpp — R expression parser, deparser, validator, Stan and CUDA emitters (see rexpr.hpp).
#include "rexpr.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace bpc {

// ------------------------------------------------------------------------------------------------
// number formatting: R prints doubles with up to 15 significant digits, dropping trailing zeros,
// and uses fixed notation unless the scientific form is strictly narrower (R_print.scipen = 0).
// This is what paste()/as.character()/deparse() apply to the literals of a rate expression
// (R/stan_utils.R:215-218) and to prior parameters (R/stan_utils.R:395-396).
std::string r_num_to_string(double x) {
  if (std::isnan(x)) return "NaN";
  if (std::isinf(x)) return x > 0 ? "Inf" : "-Inf";
  if (x == 0.0) return "0";
  char buf[64];
  // minimal number of significant digits (<= 15) that reproduces the 15-digit rounding
  std::snprintf(buf, sizeof buf, "%.14e", x);
  const double x15 = std::strtod(buf, nullptr);
  int nsig = 15;
  for (int d = 1; d <= 15; ++d) {
    std::snprintf(buf, sizeof buf, "%.*e", d - 1, x);
    if (std::strtod(buf, nullptr) == x15) { nsig = d; break; }
  }
  std::snprintf(buf, sizeof buf, "%.*e", nsig - 1, x);
  // split mantissa / exponent
  std::string s(buf);
  const size_t epos = s.find('e');
  std::string mant = s.substr(0, epos);
  const int ex = std::atoi(s.c_str() + epos + 1);
  const bool neg = mant[0] == '-';
  if (neg) mant.erase(0, 1);
  std::string digits;
  for (char ch : mant) if (ch != '.') digits.push_back(ch);
  while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
  const int nd = (int)digits.size();
  // scientific form: d[.ddd]e[+-]XX (at least two exponent digits)
  std::string sci = digits.substr(0, 1);
  if (nd > 1) sci += "." + digits.substr(1);
  char eb[16];
  std::snprintf(eb, sizeof eb, "e%c%02d", ex < 0 ? '-' : '+', std::abs(ex));
  sci += eb;
  // fixed form
  std::string fix;
  if (ex >= 0) {
    if (nd <= ex + 1) fix = digits + std::string((size_t)(ex + 1 - nd), '0');
    else fix = digits.substr(0, (size_t)ex + 1) + "." + digits.substr((size_t)ex + 1);
  } else {
    fix = "0." + std::string((size_t)(-ex - 1), '0') + digits;
  }
  const std::string& out = fix.size() <= sci.size() ? fix : sci;
  return (neg ? "-" : "") + out;
}

// ------------------------------------------------------------------------------------------------
// tokenizer
namespace {

enum class Tok { Num, Str, Id, Op, LParen, RParen, LBrack, RBrack, LDBrack, Comma, End };
struct Token { Tok t; std::string s; double v = 0; bool isint = false; };

struct Lexer {
  const std::string& src;
  size_t i = 0;
  int brack_depth_guard = 0;
  explicit Lexer(const std::string& s) : src(s) {}
  [[noreturn]] void fail() const { throw ExprError("Invalid expression: " + src); }

  Token next() {
    while (i < src.size() && (src[i] == ' ' || src[i] == '\t' || src[i] == '\n' || src[i] == '\r')) ++i;
    Token k;
    if (i >= src.size()) { k.t = Tok::End; return k; }
    const char c = src[i];
    if (std::isdigit((unsigned char)c) || (c == '.' && i + 1 < src.size() && std::isdigit((unsigned char)src[i + 1]))) {
      size_t j = i;
      if (c == '0' && j + 1 < src.size() && (src[j + 1] == 'x' || src[j + 1] == 'X')) {
        j += 2;
        while (j < src.size() && std::isxdigit((unsigned char)src[j])) ++j;
      } else {
        while (j < src.size() && std::isdigit((unsigned char)src[j])) ++j;
        if (j < src.size() && src[j] == '.') { ++j; while (j < src.size() && std::isdigit((unsigned char)src[j])) ++j; }
        if (j < src.size() && (src[j] == 'e' || src[j] == 'E')) {
          size_t q = j + 1;
          if (q < src.size() && (src[q] == '+' || src[q] == '-')) ++q;
          if (q < src.size() && std::isdigit((unsigned char)src[q])) { j = q; while (j < src.size() && std::isdigit((unsigned char)src[j])) ++j; }
        }
      }
      k.t = Tok::Num;
      k.s = src.substr(i, j - i);
      k.v = std::strtod(k.s.c_str(), nullptr);
      if (j < src.size() && src[j] == 'L') { k.isint = true; ++j; }
      else if (j < src.size() && src[j] == 'i') fail();   // complex literals are never valid rates
      i = j;
      return k;
    }
    if (std::isalpha((unsigned char)c) || c == '.' || c == '_') {
      size_t j = i;
      while (j < src.size() && (std::isalnum((unsigned char)src[j]) || src[j] == '.' || src[j] == '_')) ++j;
      k.t = Tok::Id; k.s = src.substr(i, j - i); i = j; return k;
    }
    if (c == '"' || c == '\'') {
      size_t j = i + 1; std::string out;
      while (j < src.size() && src[j] != c) {
        if (src[j] == '\\' && j + 1 < src.size()) {
          const char e = src[j + 1];
          out.push_back(e == 'n' ? '\n' : e == 't' ? '\t' : e);
          j += 2;
        } else out.push_back(src[j++]);
      }
      if (j >= src.size()) fail();
      k.t = Tok::Str; k.s = out; i = j + 1; return k;
    }
    if (c == '`') {
      size_t j = src.find('`', i + 1);
      if (j == std::string::npos) fail();
      k.t = Tok::Id; k.s = src.substr(i + 1, j - i - 1); i = j + 1; return k;
    }
    auto two = [&](const char* p) { return src.compare(i, 2, p) == 0; };
    if (c == '(') { ++i; k.t = Tok::LParen; return k; }
    if (c == ')') { ++i; k.t = Tok::RParen; return k; }
    if (c == ',') { ++i; k.t = Tok::Comma; return k; }
    if (c == '[') { if (two("[[")) { i += 2; k.t = Tok::LDBrack; } else { ++i; k.t = Tok::LBrack; } return k; }
    if (c == ']') { ++i; k.t = Tok::RBrack; return k; }
    if (c == '%') {
      size_t j = src.find('%', i + 1);
      if (j == std::string::npos) fail();
      k.t = Tok::Op; k.s = src.substr(i, j - i + 1); i = j + 1; return k;
    }
    static const char* ops2[] = {"<=", ">=", "==", "!=", "&&", "||", "<-", "->", "|>", "**"};
    for (const char* p : ops2) if (two(p)) { i += 2; k.t = Tok::Op; k.s = (std::strcmp(p, "**") == 0) ? "^" : p; return k; }
    if (std::strchr("+-*/^<>!&|:~=$@?", c)) { ++i; k.t = Tok::Op; k.s = std::string(1, c); return k; }
    fail();
  }
};

// binary operator table: left binding power, right-associative?
struct BinInfo { int lbp; bool right; };
bool bin_info(const std::string& op, BinInfo& out) {
  if (op == "^") { out = {140, true}; return true; }
  if (op == ":") { out = {120, false}; return true; }
  if (op.size() >= 2 && op.front() == '%' && op.back() == '%') { out = {110, false}; return true; }
  if (op == "|>") { out = {110, false}; return true; }
  if (op == "*" || op == "/") { out = {100, false}; return true; }
  if (op == "+" || op == "-") { out = {90, false}; return true; }
  if (op == "<" || op == ">" || op == "<=" || op == ">=" || op == "==" || op == "!=") { out = {80, false}; return true; }
  if (op == "&" || op == "&&") { out = {60, false}; return true; }
  if (op == "|" || op == "||") { out = {50, false}; return true; }
  if (op == "~") { out = {40, false}; return true; }
  if (op == "->") { out = {30, false}; return true; }
  if (op == "<-") { out = {20, true}; return true; }
  if (op == "=") { out = {10, true}; return true; }
  if (op == "?") { out = {5, false}; return true; }
  return false;
}
constexpr int UNARY_MINUS_BP = 130;   // below ^, above : and * /   (R: -2^2 == -4, -1:3 == (-1):3)
constexpr int UNARY_NOT_BP = 70;

struct Parser {
  Lexer lx;
  Token cur;
  explicit Parser(const std::string& s) : lx(s) { cur = lx.next(); }
  void advance() { cur = lx.next(); }
  [[noreturn]] void fail() const { lx.fail(); }

  NodePtr mk(NodeKind k) { auto n = std::make_unique<Node>(); n->kind = k; return n; }

  std::vector<NodePtr> args(Tok close) {
    std::vector<NodePtr> out;
    if (cur.t == close) { advance(); return out; }
    for (;;) {
      if (cur.t == Tok::Comma || cur.t == close) out.push_back(mk(NodeKind::Empty));
      else out.push_back(expr(0));
      if (cur.t == Tok::Comma) { advance(); if (cur.t == close) { out.push_back(mk(NodeKind::Empty)); advance(); return out; } continue; }
      if (cur.t == close) { advance(); return out; }
      fail();
    }
  }

  NodePtr primary() {
    NodePtr n;
    if (cur.t == Tok::Num) { n = mk(NodeKind::Num); n->num = cur.v; n->int_suffix = cur.isint; n->text = cur.s; advance(); }
    else if (cur.t == Tok::Str) { n = mk(NodeKind::Str); n->text = cur.s; advance(); }
    else if (cur.t == Tok::Id) {
      if (cur.s == "function" || cur.s == "if" || cur.s == "for" || cur.s == "while" || cur.s == "repeat") fail();
      n = mk(NodeKind::Sym); n->text = cur.s; advance();
    }
    else if (cur.t == Tok::LParen) {
      advance();
      n = mk(NodeKind::Paren);
      n->kids.push_back(expr(0));
      if (cur.t != Tok::RParen) fail();
      advance();
    }
    else if (cur.t == Tok::Op && (cur.s == "-" || cur.s == "+" || cur.s == "!" || cur.s == "~" || cur.s == "?")) {
      const std::string op = cur.s;
      advance();
      n = mk(NodeKind::Unary); n->text = op;
      n->kids.push_back(expr(op == "!" ? UNARY_NOT_BP : (op == "~" ? 40 : (op == "?" ? 5 : UNARY_MINUS_BP))));
      return n;   // postfix operators were consumed by the operand at its own precedence
    }
    else fail();
    // postfix: call, [, [[, $, @
    for (;;) {
      if (cur.t == Tok::LParen) {
        advance();
        auto c = mk(NodeKind::Call);
        c->kids.push_back(std::move(n));
        for (auto& a : args(Tok::RParen)) c->kids.push_back(std::move(a));
        n = std::move(c);
      } else if (cur.t == Tok::LBrack) {
        advance();
        auto c = mk(NodeKind::Index); c->text = "[";
        c->kids.push_back(std::move(n));
        for (auto& a : args(Tok::RBrack)) c->kids.push_back(std::move(a));
        n = std::move(c);
      } else if (cur.t == Tok::LDBrack) {
        advance();
        auto c = mk(NodeKind::Index); c->text = "[[";
        c->kids.push_back(std::move(n));
        for (auto& a : args(Tok::RBrack)) c->kids.push_back(std::move(a));
        if (cur.t != Tok::RBrack) fail();
        advance();
        n = std::move(c);
      } else if (cur.t == Tok::Op && (cur.s == "$" || cur.s == "@")) {
        const std::string op = cur.s;
        advance();
        if (cur.t != Tok::Id && cur.t != Tok::Str) fail();
        auto c = mk(NodeKind::Binary); c->text = op;
        auto r = mk(NodeKind::Sym); r->text = cur.s; advance();
        c->kids.push_back(std::move(n)); c->kids.push_back(std::move(r));
        n = std::move(c);
      } else break;
    }
    return n;
  }

  NodePtr expr(int min_bp) {
    NodePtr lhs = primary();
    for (;;) {
      if (cur.t != Tok::Op) break;
      BinInfo bi;
      if (!bin_info(cur.s, bi) || bi.lbp <= min_bp) break;
      const std::string op = cur.s;
      advance();
      // right-associative operators re-enter at their own level; the exponent of ^ may carry a unary sign (2^-1)
      NodePtr rhs = expr(bi.right ? bi.lbp - 1 : bi.lbp);
      auto b = mk(NodeKind::Binary); b->text = op;
      b->kids.push_back(std::move(lhs)); b->kids.push_back(std::move(rhs));
      lhs = std::move(b);
    }
    return lhs;
  }
};

bool all_digits(const std::string& s) {
  if (s.empty()) return false;
  for (char c : s) if (!std::isdigit((unsigned char)c)) return false;
  return true;
}

std::string deparse_string(const std::string& s) {
  std::string o = "\"";
  for (char c : s) {
    if (c == '"' || c == '\\') { o.push_back('\\'); o.push_back(c); }
    else if (c == '\n') o += "\\n";
    else if (c == '\t') o += "\\t";
    else o.push_back(c);
  }
  return o + "\"";
}

}  // namespace

NodePtr parse_r(const std::string& src) {
  Parser p(src);
  if (p.cur.t == Tok::End) p.fail();
  NodePtr n = p.expr(0);
  if (p.cur.t != Tok::End) p.fail();
  return n;
}

std::string deparse(const Node& n) {
  switch (n.kind) {
    case NodeKind::Empty: return "";
    case NodeKind::Num: return r_num_to_string(n.num) + (n.int_suffix ? "L" : "");
    case NodeKind::Str: return deparse_string(n.text);
    case NodeKind::Sym: return n.text;
    case NodeKind::Paren: return "(" + deparse(*n.kids[0]) + ")";
    case NodeKind::Unary: return n.text + deparse(*n.kids[0]);
    case NodeKind::Binary: {
      const std::string& op = n.text;
      const bool tight = (op == "^" || op == "/" || op == ":" || op == "$" || op == "@");
      return deparse(*n.kids[0]) + (tight ? op : " " + op + " ") + deparse(*n.kids[1]);
    }
    case NodeKind::Call:
    case NodeKind::Index: {
      std::string o = deparse(*n.kids[0]);
      o += n.kind == NodeKind::Call ? "(" : n.text;
      for (size_t i = 1; i < n.kids.size(); ++i) { if (i > 1) o += ", "; o += deparse(*n.kids[i]); }
      o += n.kind == NodeKind::Call ? ")" : (n.text == "[" ? "]" : "]]");
      return o;
    }
  }
  return "";
}

namespace {

bool is_arith(const std::string& op) { return op == "+" || op == "-" || op == "/" || op == "*" || op == "^"; }

// the `[` branch of exp_to_stan / check_valid: R/stan_utils.R:249-275, 334-360.  Returns 'c' or 'x' and the index.
char classify_index(const Node& n, int max_params, int max_dep, int& num) {
  if (n.text == "[" && n.kids.size() == 2 && n.kids[0]->kind == NodeKind::Sym && n.kids[1]->kind == NodeKind::Num) {
    const std::string idx = deparse(*n.kids[1]);
    const std::string& nm = n.kids[0]->text;
    if (all_digits(idx) && (nm == "c" || nm == "x")) {
      num = (int)std::strtol(idx.c_str(), nullptr, 10);
      const std::string s = nm + "[" + idx + "]";
      if (nm == "c") {
        if (num > max_params || num <= 0) throw ExprError("Parameter " + s + " goes beyond the number of parameters specified.");
        return 'c';
      }
      if (num > max_dep || num <= 0) throw ExprError("Variable " + s + " goes beyond the number of dependent variables specified.");
      return 'x';
    }
  }
  throw ExprError("Invalid expression: " + deparse(n));
}

}  // namespace

void check_valid(const Node& n, int max_params, int max_dep) {
  switch (n.kind) {
    case NodeKind::Num: return;
    case NodeKind::Paren: check_valid(*n.kids[0], max_params, max_dep); return;
    case NodeKind::Unary:
      if (n.text == "+" || n.text == "-") { check_valid(*n.kids[0], max_params, max_dep); return; }
      break;
    case NodeKind::Binary:
      if (is_arith(n.text)) { check_valid(*n.kids[0], max_params, max_dep); check_valid(*n.kids[1], max_params, max_dep); return; }
      break;
    case NodeKind::Call:
      if (n.kids[0]->kind == NodeKind::Sym && (n.kids[0]->text == "exp" || n.kids[0]->text == "log")) {
        if (n.kids.size() > 2) throw ExprError("log and exp take only one parameter");
        if (n.kids.size() == 2 && n.kids[1]->kind != NodeKind::Empty) { check_valid(*n.kids[1], max_params, max_dep); return; }
      }
      break;
    case NodeKind::Index: { int num; classify_index(n, max_params, max_dep, num); return; }
    default: break;
  }
  throw ExprError("Invalid expression: " + deparse(n));
}

std::string to_stan(const Node& n, int max_params, int max_dep) {
  switch (n.kind) {
    case NodeKind::Num: return r_num_to_string(n.num);
    case NodeKind::Paren: return to_stan(*n.kids[0], max_params, max_dep);
    case NodeKind::Unary:
      if (n.text == "+" || n.text == "-") return n.text + "(" + to_stan(*n.kids[0], max_params, max_dep) + ")";
      break;
    case NodeKind::Binary:
      if (is_arith(n.text))
        return "( " + to_stan(*n.kids[0], max_params, max_dep) + " ) " + n.text + " ( " + to_stan(*n.kids[1], max_params, max_dep) + " )";
      break;
    case NodeKind::Call:
      if (n.kids[0]->kind == NodeKind::Sym && (n.kids[0]->text == "exp" || n.kids[0]->text == "log")) {
        if (n.kids.size() > 2) throw ExprError("log and exp take only one parameter");
        if (n.kids.size() == 2 && n.kids[1]->kind != NodeKind::Empty)
          return n.kids[0]->text + "(" + to_stan(*n.kids[1], max_params, max_dep) + ")";
      }
      break;
    case NodeKind::Index: {
      int num;
      const char k = classify_index(n, max_params, max_dep, num);
      char b[64];
      if (k == 'c') std::snprintf(b, sizeof b, "Theta%d", num);
      else std::snprintf(b, sizeof b, "function_var[i,%d]", num);
      return b;
    }
    default: break;
  }
  throw ExprError("Invalid expression: " + deparse(n));
}

// ------------------------------------------------------------------------------------------------
// lowering to CUDA
namespace {

struct Val {
  enum Kind { IntConst, RealConst, Var } kind = RealConst;
  long long i = 0;
  double d = 0.0;
  std::string name;     // Var: C expression naming the value (temp, c[k] or x[k])
  std::string adj;      // Var depending on c: C lvalue accumulating its adjoint ("" if none)
  bool depc = false, depx = false;
};

std::string lit(double v) {
  if (std::isnan(v)) return "(0.0/0.0)";
  if (std::isinf(v)) return v > 0 ? "(1.0/0.0)" : "(-1.0/0.0)";
  char b[64];
  std::snprintf(b, sizeof b, "%.17g", v);
  std::string s(b);
  if (s.find_first_of(".eEn") == std::string::npos) s += ".0";
  if (v < 0) s = "(" + s + ")";
  return s;
}

struct Lowerer {
  int e, max_params, max_dep, ntemp = 0;
  std::vector<std::string> fwd;      // forward statements
  std::vector<std::string> adjdecl;  // adjoint declarations
  std::vector<std::string> rev;      // reverse statements, in forward order (emitted reversed)
  std::vector<std::string> warnings;
  bool uses_x = false;

  std::string cexpr(const Val& v) const {
    if (v.kind == Val::IntConst) return lit((double)v.i);
    if (v.kind == Val::RealConst) return lit(v.d);
    return v.name;
  }
  double cval(const Val& v) const { return v.kind == Val::IntConst ? (double)v.i : v.d; }

  Val temp(const std::string& rhs, bool depc, bool depx) {
    Val v; v.kind = Val::Var; v.depc = depc; v.depx = depx;
    char b[48]; std::snprintf(b, sizeof b, "e%d_t%d", e, ntemp++);
    v.name = b;
    fwd.push_back("const real " + v.name + " = " + rhs + ";");
    if (depc) { v.adj = v.name + "_b"; adjdecl.push_back("real " + v.adj + " = real(0.0);"); }
    return v;
  }

  Val lower(const Node& n) {
    switch (n.kind) {
      case NodeKind::Num: {
        // the literal reaches Stan through paste(): 15 significant digits; no '.'/'e' => Stan int
        const std::string s = r_num_to_string(n.num);
        Val v;
        if (s.find_first_of(".eEIN") == std::string::npos && std::fabs(n.num) < 9e15) { v.kind = Val::IntConst; v.i = std::strtoll(s.c_str(), nullptr, 10); }
        else { v.kind = Val::RealConst; v.d = std::strtod(s.c_str(), nullptr); }
        return v;
      }
      case NodeKind::Paren: return lower(*n.kids[0]);
      case NodeKind::Index: {
        int num;
        const char k = classify_index(n, max_params, max_dep, num);
        Val v; v.kind = Val::Var;
        char b[48];
        if (k == 'c') { std::snprintf(b, sizeof b, "c[%d]", num - 1); v.name = b; std::snprintf(b, sizeof b, "cb[%d]", num - 1); v.adj = b; v.depc = true; }
        else { std::snprintf(b, sizeof b, "x[%d]", num - 1); v.name = b; v.depx = true; uses_x = true; }
        return v;
      }
      case NodeKind::Unary: {
        if (n.text != "+" && n.text != "-") break;
        Val a = lower(*n.kids[0]);
        if (n.text == "+") return a;
        if (a.kind == Val::IntConst) { a.i = -a.i; return a; }
        if (a.kind == Val::RealConst) { a.d = -a.d; return a; }
        Val z = temp("-" + a.name, a.depc, a.depx);
        if (a.depc) rev.push_back(a.adj + " -= " + z.adj + ";");
        return z;
      }
      case NodeKind::Call: {
        if (!(n.kids[0]->kind == NodeKind::Sym && (n.kids[0]->text == "exp" || n.kids[0]->text == "log"))) break;
        if (n.kids.size() > 2) throw ExprError("log and exp take only one parameter");
        if (n.kids.size() != 2 || n.kids[1]->kind == NodeKind::Empty) break;
        const bool is_exp = n.kids[0]->text == "exp";
        Val a = lower(*n.kids[1]);
        if (a.kind != Val::Var) { Val v; v.kind = Val::RealConst; v.d = is_exp ? std::exp(cval(a)) : std::log(cval(a)); return v; }
        Val z = temp(std::string(is_exp ? "exp(" : "log(") + a.name + ")", a.depc, a.depx);
        if (a.depc) rev.push_back(a.adj + " += " + z.adj + (is_exp ? " * " + z.name : " / " + a.name) + ";");
        return z;
      }
      case NodeKind::Binary: {
        if (!is_arith(n.text)) break;
        Val a = lower(*n.kids[0]);
        Val b = lower(*n.kids[1]);
        const char op = n.text[0];
        if (a.kind != Val::Var && b.kind != Val::Var) {
          Val v;
          if (a.kind == Val::IntConst && b.kind == Val::IntConst && op != '^') {
            v.kind = Val::IntConst;
            if (op == '+') v.i = a.i + b.i;
            else if (op == '-') v.i = a.i - b.i;
            else if (op == '*') v.i = a.i * b.i;
            else {
              if (b.i == 0) { v.kind = Val::RealConst; v.d = std::nan(""); warnings.push_back("integer division by zero in rate expression of event " + std::to_string(e + 1)); return v; }
              v.i = a.i / b.i;   // Stan int division truncates; stanc warns
              if (a.i % b.i != 0)
                warnings.push_back("Stan integer division: " + std::to_string(a.i) + " / " + std::to_string(b.i) + " evaluates to " + std::to_string(v.i) +
                                   " in rate expression of event " + std::to_string(e + 1) + " (R pastes whole numbers without a decimal point, so Stan sees two ints; write the quotient as a decimal literal)");
            }
            return v;
          }
          v.kind = Val::RealConst;
          const double x = cval(a), y = cval(b);
          v.d = op == '+' ? x + y : op == '-' ? x - y : op == '*' ? x * y : op == '/' ? x / y : std::pow(x, y);
          return v;
        }
        const bool depc = a.depc || b.depc, depx = a.depx || b.depx;
        const std::string A = cexpr(a), B = cexpr(b);
        if (op == '^') {
          Val z = temp("pow(" + A + ", " + B + ")", depc, depx);
          if (a.depc) rev.push_back(a.adj + " += " + z.adj + " * " + B + " * pow(" + A + ", " + B + " - real(1.0));");
          if (b.depc) rev.push_back(b.adj + " += " + z.adj + " * " + z.name + " * log(" + A + ");");
          return z;
        }
        Val z = temp(A + " " + op + " " + B, depc, depx);
        if (op == '+') { if (a.depc) rev.push_back(a.adj + " += " + z.adj + ";"); if (b.depc) rev.push_back(b.adj + " += " + z.adj + ";"); }
        else if (op == '-') { if (a.depc) rev.push_back(a.adj + " += " + z.adj + ";"); if (b.depc) rev.push_back(b.adj + " -= " + z.adj + ";"); }
        else if (op == '*') { if (a.depc) rev.push_back(a.adj + " += " + z.adj + " * " + B + ";"); if (b.depc) rev.push_back(b.adj + " += " + z.adj + " * " + A + ";"); }
        else { if (a.depc) rev.push_back(a.adj + " += " + z.adj + " / " + B + ";"); if (b.depc) rev.push_back(b.adj + " -= " + z.adj + " * " + z.name + " / " + B + ";"); }
        return z;
      }
      default: break;
    }
    throw ExprError("Invalid expression: " + deparse(n));
  }
};

}  // namespace

RateProgram lower_to_cuda(const Node& n, int e, int max_params, int max_dep) {
  Lowerer lw{e, max_params, max_dep};
  Val top = lw.lower(n);
  RateProgram p;
  p.uses_x = lw.uses_x;
  p.warnings = lw.warnings;
  const std::string re = "r[" + std::to_string(e) + "]", rbe = "rb[" + std::to_string(e) + "]";
  std::string f;
  for (auto& s : lw.fwd) f += "  " + s + "\n";
  p.value_code = "  {\n" + f + "  " + re + " = " + lw.cexpr(top) + ";\n  }\n";
  if (top.kind == Val::Var && top.depc) {
    std::string v = "  {\n" + f;
    for (auto& s : lw.adjdecl) v += "  " + s + "\n";
    // seed: adjoint of the top value is rb[e]
    v += "  " + top.adj + " += " + rbe + ";\n";
    for (auto it = lw.rev.rbegin(); it != lw.rev.rend(); ++it) v += "  " + *it + "\n";
    p.vjp_code = v + "  }\n";
  } else {
    p.vjp_code = "";
  }
  return p;
}

}  // namespace bpc

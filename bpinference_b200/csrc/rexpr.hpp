// rexpr.hpp — front end for the user's rate expressions.
//
// The reference keeps each rate as a parsed R call (R/bp_model.R:46-50) and walks it twice:
// check_valid (R/stan_utils.R:289-364) and exp_to_stan (R/stan_utils.R:205-281, emits Stan text).
// There is no R here, so this file parses the R expression grammar itself, reproduces R's
// deparse() spelling (the reference's error messages quote it, tests/testthat/
// test_expression_parsing.R:4-23), performs the same validation walk with the same messages,
// and has two emitters: Stan text (what exp_to_stan returns) and CUDA device code for the rate
// and its reverse-mode derivative (the sibling the north star asks for).
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace bpc {

struct ExprError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

enum class NodeKind { Num, Str, Sym, Call, Index, Unary, Binary, Paren, Empty };

struct Node {
  NodeKind kind = NodeKind::Empty;
  double num = 0.0;          // Num
  bool int_suffix = false;   // 1L
  std::string text;          // Str contents / Sym name / operator spelling / Index: "[" or "[["
  std::vector<std::unique_ptr<Node>> kids;  // Call: kids[0] = function, rest = arguments
};
using NodePtr = std::unique_ptr<Node>;

// R's as.character / deparse of a double (15 significant digits, fixed unless scientific is narrower)
std::string r_num_to_string(double x);

NodePtr parse_r(const std::string& src);       // throws ExprError("Invalid expression: <src>") on syntax errors
std::string deparse(const Node& n);            // R's deparse()
void check_valid(const Node& n, int max_params, int max_dep);             // throws ExprError with the reference's text
std::string to_stan(const Node& n, int max_params, int max_dep);          // exp_to_stan

// ---- lowered form used by the CUDA emitter ------------------------------------------------------
// A rate is lowered to straight-line SSA over doubles.  Stan's typing of literals is kept: a literal that
// R pastes without '.'/'e' is a Stan int, int (+,-,*,/) int stays int with C++ truncating division
// (SURVEY.md §7 quirk list), everything else is real.  Int sub-expressions are folded to constants.
struct RateProgram {
  std::string value_code;   // statements computing r[e] from c[], x[]
  std::string vjp_code;     // statements accumulating cb[] += d r[e]/d c * rb[e]
  bool uses_x = false;      // depends on a dependent variable x[k]
  std::vector<std::string> warnings;
};
// emits code for event `e` (0-based); variable names are prefixed so programs can share a scope
RateProgram lower_to_cuda(const Node& n, int e, int max_params, int max_dep);

}  // namespace bpc

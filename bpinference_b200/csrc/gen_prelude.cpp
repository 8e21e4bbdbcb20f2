// gen_prelude.cpp — build-time tool: prints the generated translation unit for the default 3-type model
// (see default_model.cu).
#include <cstdio>

#include "codegen.hpp"

int main() {
  bpc::ModelSpec m;
  m.ntypes = 3; m.nevents = 9; m.nparams = 9; m.ndep = 0; m.kind = 0;
  m.e_mat.assign(27, 0.0);
  for (int i = 0; i < 3; ++i) {
    m.e_mat[3 * i + 9 * i] = 2.0;                  // division: two of type i
    m.e_mat[3 * i + 1 + 9 * ((i + 1) % 3)] = 1.0;  // transition to type i+1
    for (int q = 0; q < 3; ++q) m.p_vec.push_back(i + 1);
  }
  for (int k = 0; k < 9; ++k) {
    m.func_deps.push_back("c[" + std::to_string(k + 1) + "]");
    bpc::PriorSpec p;
    p.name = "normal"; p.nparams = 2; p.params[0] = 0.0; p.params[1] = 0.25; p.has_bounds = true; p.lower = 0.0; p.upper = 5.0;
    m.priors.push_back(p);
  }
  std::fputs(bpc::generate_cuda(m).c_str(), stdout);
  return 0;
}

// codegen.hpp — turns a validated model description into the CUDA translation unit that NVRTC compiles.
// The sibling of exp_to_stan / generate() (R/stan_utils.R:130-281): same inputs (model, priors, template
// switch), CUDA device code instead of Stan text.
#pragma once
#include <string>
#include <vector>

namespace bpc {

struct PriorSpec {
  std::string name;
  int id = -1;
  int nparams = 0;
  double params[3] = {0, 0, 0};
  bool has_bounds = false;
  double lower = 0, upper = 0;
};

struct ModelSpec {
  int ntypes = 0, nevents = 0, nparams = 0, ndep = 0, kind = 0;
  std::vector<double> e_mat;            // nevents x ntypes column-major
  std::vector<int> p_vec;               // 1-based
  std::vector<std::string> func_deps;
  std::vector<PriorSpec> priors;
  // derived
  int dim = 0;        // nparams (+1 for the noise template)
  int kdep = 1;       // stored columns of dependent variables (>= 1)
  bool xdep = false;  // some rate reads x[k]
  int threads = 128;  // CTA size of the per-observation kernels
  int mcap = 23;      // the series has at most mcap + 1 applications of L per sub-step
  int wcap = 31;      // W path: at most wcap + 1 applications of L (no ring, so longer single-step series are affordable)
  int mslot = 12;     // Taylor terms kept per thread in shared memory
  int minblocks = 1;  // CTAs per SM the shared-memory footprint allows
  double theta_cap = 0;
  size_t w_smem = 0;        // dynamic shared memory of bp_fusedw
  size_t c_smem = 0;        // dynamic shared memory of bp_fusedc
  int w_post_threads = 128; // CTA size of bp_wpost
  std::vector<bool> a_nz, h_nz;   // structural non-zeros of A [n*n] and H [n*s]
  std::vector<std::string> warnings;
};

// ALLOWED_DISTRIBUTIONS (R/sysdata.rda): id, parameter count and positivity constraints; throws
// std::runtime_error with the reference's message (R/stan_utils.R:418-459) when the prior is invalid
void validate_prior(PriorSpec& p);
int prior_id(const std::string& name);   // -1 when unknown

// validates shapes (messages of R/bp_model.R:24-68 where they apply to a C caller), parses and checks the
// rate expressions, fills the derived fields and returns the CUDA source
std::string generate_cuda(ModelSpec& m);

// Taylor degree table shared with the kernels: th[d]^d / d! = 2^-55
double taylor_theta(int d);

}  // namespace bpc

// sampler.cu — placeholder until the device-resident HMC driver lands (next commit).
#include "bpcuda.h"
#include "bpcuda_internal.hpp"
using namespace bpc;
extern "C" {
int bpc_sampler_create(bpc_model*, bpc_data*, const bpc_sample_args*, bpc_sampler**) { return fail(BPC_E_UNSUPPORTED, "sampler not built yet"); }
void bpc_sampler_free(bpc_sampler*) {}
int bpc_sampler_run(bpc_sampler*, int, int, int*, int*) { return fail(BPC_E_UNSUPPORTED, "sampler not built yet"); }
int bpc_sampler_stats_dev(bpc_sampler*, double**, int*) { return fail(BPC_E_UNSUPPORTED, "sampler not built yet"); }
int bpc_sampler_apply_stats(bpc_sampler*) { return fail(BPC_E_UNSUPPORTED, "sampler not built yet"); }
int bpc_sampler_draws(bpc_sampler*, double*, double*, int*, int*, int*, double*, double*) { return fail(BPC_E_UNSUPPORTED, "sampler not built yet"); }
long long bpc_sampler_total_leapfrogs(const bpc_sampler*) { return 0; }
int bpc_sample(bpc_model*, bpc_data*, const bpc_sample_args*, double*, double*, int*, int*, int*, double*, double*) { return fail(BPC_E_UNSUPPORTED, "sampler not built yet"); }
int bpc_simulate(int, int, const double*, const int*, const double*, const double*, const double*, int, int, unsigned long long, int, double*) { return fail(BPC_E_UNSUPPORTED, "simulator not built yet"); }
}

"""bpinference_b200 — B200-native engine for bpinference's posterior log-density path.

Host-side mirror of the reference's R interface (same names, argument meaning and error text) on top of
libbpcuda's C ABI (include/bpcuda.h).  See DESIGN.md for the path, its boundary and the kernels.
"""
from .model import BpModel, bp_model, bp_model_simple_birth_death, check_valid, deparse, exp_to_stan, new_bp_model, validate_bp_model
from .priors import ALLOWED_DISTRIBUTIONS, PriorDist, new_prior_dist, parse_prior, prior_dist, validate_prior_dist
from .stan_data import create_stan_data, stan_data_from_simulation, uniform_initialize
from .simulation import bp, bp_cuda, bpsims
from .moments import calculate_moments, template_moments
from .engine import BpCudaModel, StanData, fp64_peak, fp64_tensor_peak, stan_model
from .fit import BpFit, check_div, check_exceeded_treedepth, check_rhat, sampling, sampling_multi

__all__ = [
    "BpModel", "bp_model", "bp_model_simple_birth_death", "new_bp_model", "validate_bp_model", "check_valid", "deparse", "exp_to_stan",
    "ALLOWED_DISTRIBUTIONS", "PriorDist", "prior_dist", "new_prior_dist", "validate_prior_dist", "parse_prior",
    "create_stan_data", "stan_data_from_simulation", "uniform_initialize", "bp", "bp_cuda", "bpsims", "calculate_moments", "template_moments",
    "BpCudaModel", "StanData", "stan_model", "fp64_peak", "fp64_tensor_peak", "BpFit", "sampling", "sampling_multi", "check_div", "check_exceeded_treedepth", "check_rhat",
]

"""unbiased_pimh_b200 -- B200-native (sm_100a CUDA) bootstrap-particle-filter hot path of the R package CoupledPIMH
(particlemontecarlo/unbiased_pimh), behind the reference's own API names.  See DESIGN.md and INTEGRATION.md.

The compute library is libcpimh.so (C ABI: include/cpimh.h); this package is the thin host-side mirror of the
reference's R interface.  Importing it does not need a GPU; every compute call does, and fails loudly without one.
"""
from ._lib import CpimhError, lib  # noqa: F401
from .rng import set_seed  # noqa: F401
from .models import Model, get_gss, get_ar, get_levy_sv, get_stochastic_kinetic, get_lorenz, create_A  # noqa: F401
from .resampling import (multinomial_resampling_n, multinomial_resampling_n_, systematic_resampling_n, systematic_resampling_n_,  # noqa: F401
                         coupled_multinomial_resampling_n_, indexmatching_cpp, CR_multinomial, CR_systematic, CR_indexmatching,
                         normalize_weight, sorted_uniforms)
from .filters import CPF, CPF_coupled, CPF_RB, CPF_RB_coupled, ccpf_batch, pf, pf_batch, particlefilter, cpf_batch, PfBatch, pf_get_weighted_average  # noqa: F401
from .pimh import (get_pimh_kernels, coupled_pimh_chains, H_bar, BC, rheeglynn_estimator, coupled_pimh_chains_batch, ChainBatch,  # noqa: F401
                   unbiased_estimator_sharded, shard_range, combine_sums, get_cpf_kernels, coupled_ccpf_chains,
                   coupled_ccpf_chains_batch, rheeglynn_estimator_RB, unbiased_ccpf_estimator_sharded)
from .tree import Tree  # noqa: F401
from .mvnorm import (fast_dmvnorm, fast_dmvnorm_transpose, fast_dmvnorm_transpose_cholesky, fast_rmvnorm, fast_rmvnorm_transpose,  # noqa: F401
                     fast_rmvnorm_transpose_cholesky, lorenz_transition)
from .rdata import read_rdata, write_rdata  # noqa: F401

__version__ = "0.1.0"

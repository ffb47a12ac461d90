"""pgbp_b200 -- B200-native batched Gaussian belief propagation.

Drop-in for the message-passing hot path of PhyloGaussianBeliefProp.jl
(calibrate! / propagate_belief! / integratebelief! / factored_energy /
assignfactors! / regularizebeliefs_*), batched over trait replicates and
parameter vectors, running as hand-written sm_100a CUDA kernels behind the
C ABI in include/pgbp_b200.h.  See DESIGN.md and INTEGRATION.md.
"""
from . import _lib
from ._lib import Library, PgbpError, default_library
from .api import (BatchedClusterGraphBelief, ClusterGraphPlan, assignfactors, bm_params, calibrate,
                  factored_energy, families_table, free_energy, init_beliefs_reset_fromfactors,
                  init_factors_frombeliefs, init_factors_frommodel, init_messagecalibrationflags_reset,
                  integratebelief, propagate_1traversal_postorder, propagate_1traversal_preorder,
                  propagate_belief, regularizebeliefs_bycluster, regularizebeliefs_bynodesubtree,
                  regularizebeliefs_onschedule, scopeindex)

from .drivers import calibrate_exact_cliquetree, calibrate_optimize_cliquetree, calibrate_optimize_clustergraph

# spellings used by the reference revision named in BASELINE.json's north_star
init_beliefs_allocate = ClusterGraphPlan.from_beliefs
__all__ = [n for n in dir() if not n.startswith("_")]

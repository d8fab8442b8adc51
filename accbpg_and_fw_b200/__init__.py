"""accbpg hot path on B200: GPU-resident drop-ins behind accbpg's own Python API.

    import accbpg_and_fw_b200 as accbpg
    f, h, L, x0 = accbpg.D_opt_design(80, 200)
    x, F, Ls, T = accbpg.BPG(f, h, L, x0, maxitrs=1000, verbskip=100)

Everything numerical runs in hand-written sm_100a kernels reached through the C ABI in
include/accbpg_b200.h (libaccbpg_b200.so).  Importing this package fails if that library has
not been built; there is no CPU fallback.
"""
from . import _native                                         # noqa: F401  (raises if the library is missing)
from .objectives import RSmoothFunction, DOptimalObj, SparseDOptimalObj, PoissonRegression, KLdivRegression
from .bregman import (LegendreFunction, BurgEntropy, BurgEntropyL1, BurgEntropyL2, BurgEntropySimplex,
                      ShannonEntropy, ShannonEntropyL1, ShannonEntropySimplex, SquaredL2Norm)
from .lmo import (lmo_simplex, lmo_matrix_simplex, lmo_l2_ball, lmo_l2_ball_positive_orthant, lmo_linf_ball,
                  lmo_matrix_box, lmo_nuclear_norm_ball)
from .drivers import BPG, ABPG, ABPG_expo, ABPG_gain, ABDA, solve_theta
from .drivers_fw import (FW_alg_div_step, FW_alg_descent_step, FW_alg_L0_L1_shortest_step,
                         FW_l0l1_log_and_linear_step, FW_l0l1_log_only)
from .dopt_fw import D_opt_FW, D_opt_FW_away
from .problems import (D_opt_libsvm, D_opt_design, D_opt_KYinit, Poisson_regrL1, Poisson_regrL2, KL_nonneg_regr,
                       load_libsvm_dense, load_libsvm_sparse)
from .dist import ColumnShard
from .runtime import Runtime

__all__ = [
    "RSmoothFunction", "DOptimalObj", "SparseDOptimalObj", "PoissonRegression", "KLdivRegression",
    "LegendreFunction", "BurgEntropy", "BurgEntropyL1", "BurgEntropyL2", "BurgEntropySimplex",
    "ShannonEntropy", "ShannonEntropyL1", "ShannonEntropySimplex", "SquaredL2Norm",
    "lmo_simplex", "lmo_matrix_simplex", "lmo_l2_ball", "lmo_l2_ball_positive_orthant", "lmo_linf_ball",
    "lmo_matrix_box", "lmo_nuclear_norm_ball",
    "BPG", "ABPG", "ABPG_expo", "ABPG_gain", "ABDA", "solve_theta",
    "FW_alg_div_step", "FW_alg_descent_step", "FW_alg_L0_L1_shortest_step", "FW_l0l1_log_and_linear_step",
    "FW_l0l1_log_only", "D_opt_FW", "D_opt_FW_away",
    "D_opt_libsvm", "D_opt_design", "D_opt_KYinit", "Poisson_regrL1", "Poisson_regrL2", "KL_nonneg_regr",
    "load_libsvm_dense", "load_libsvm_sparse", "ColumnShard", "Runtime",
]

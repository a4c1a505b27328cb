// rp_solver_wide.cu -- the T = 512 build of the relative-pose solver (one CTA per SM) for small batches.
//
// Same source as rp_solver.cu (RelativePoseEstimation_helper, RPModule/rpmodule.py:317-508); only the CTA width differs.  A
// single scan pair through the T = 128 kernel keeps one quarter of one SM busy (evaluation.py:278-284 calls the solver one
// pair at a time; the alternation of rpmodule.py:569-662 solves 32 pairs per step on 148 SMs): the pair loops -- 132 k
// pre-tests, ~10-30 k exact tests, ~250 passes over the CSR of W -- are spread over four times as many threads here.
// Exports rp_wide_solve_batch_ex / rp_wide_spectral_irls_solve, called by the entry points of rp_solver.cu for batches of at
// most rp_solver_wide_max() pairs.
#define RP_THREADS 512
#define RP_MIN_BLOCKS 1
#define RP_WIDE_TU 1
#include "rp_solver.cu"

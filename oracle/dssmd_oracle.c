/*
 * dssmd_oracle.c -- CPU oracle for the DSSMD hot path (correlation cost volume
 * + disparity warp), plain C with OpenMP over (batch, row).
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product path (dssmd_b200/) never links or calls it and has no CPU fallback.
 *
 * What it restates: the upstream repo's PyTorch implementation of
 *   build_corr            models/UtilsNet/submoudles.py:301-311
 *   CostVolume.correlation models/UtilsNet/cost_volume.py:39-47
 *   disp_warp             utils/disparity_warper.py:83-106 (== models/UtilsNet/stereo_op.py:41-63)
 *   LRwarp_error          utils/disparity_warper.py:109-115
 *   convert_disp_to_depth / depth2trans / recover_haze_images   DeBug/inference.py:57-64, 126-134, 95-112
 * and the autograd backward of each.  The arithmetic of the upstream functions
 * lives in PyTorch (aten::mul/mean, aten::grid_sampler_2d,
 * aten::upsample_bilinear2d; torch 2.11.0 in this project's image, upstream
 * does not pin a version); the op order restated here is the one spelled out in
 * SURVEY.md Appendix A.
 *
 * Parity pin: the upstream repo ships no tests or golden vectors for this
 * path.  The oracle is pinned against outputs of the upstream functions
 * themselves, imported from /root/reference on CPU by oracle/gen_golden.py and
 * committed as tests/golden/ (.npz files) (tests/test_oracle_golden.py).
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared).
 * -ffp-contract=off keeps the compiler from fusing a*b+c on its own; the one
 * place the reference's compiled kernels use an FMA is written as fma().
 */
#include <math.h>

#define REAL float
#define SUF f32
#define FMA fmaf
#define FLOOR floorf
#define FMIN fminf
#define FMAX fmaxf
#define EXP expf
#include "dssmd_oracle_impl.h"
#undef REAL
#undef SUF
#undef FMA
#undef FLOOR
#undef FMIN
#undef FMAX
#undef EXP

#define REAL double
#define SUF f64
#define FMA fma
#define FLOOR floor
#define FMIN fmin
#define FMAX fmax
#define EXP exp
#include "dssmd_oracle_impl.h"
#undef REAL
#undef SUF
#undef FMA
#undef FLOOR
#undef FMIN
#undef FMAX
#undef EXP

#ifdef _OPENMP
#include <omp.h>
int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
#else
int oracle_num_threads(void) { return 1; }
void oracle_set_num_threads(int n) { (void)n; }
#endif
int oracle_version(void) { return 1; }

// rp_solver.cu -- B200 (sm_100a) relative-pose solver: small persistent CTAs, one scan pair at a time each.
//
// Replaces the reference's RelativePoseEstimation_helper (RPModule/rpmodule.py:317-508) and the four
// fitters it dispatches to (fit_horn87 :60, fit_spectral :86, fit_irls :169, fit_irls_sm :212,
// horn87_np :17).  Nothing here is translated from the reference: the NumPy code enumerates all
// N(N-1)/2 correspondence pairs with fancy indexing, stacks 4M rows and calls ARPACK on an
// (n_s*n_t)^2 sparse matrix; this kernel
//   A. forms the n_s x n_t float32 descriptor distances with NumPy's exact summation order (so the
//      top-k index sets are bit-identical), soft-match weights, per-row top-k         (:342-375)
//   B. gathers per-correspondence geometry (float64 in the slot's workspace, float32 copy in smem)
//   C. runs a conservative float32 pre-test of the distance-consistency condition over every
//      correspondence pair and emits a compact candidate list (never rejects a pair the float64
//      test would accept: margins scale with the coordinate magnitude)               (:382-404)
//   D. per candidate: exact float64 distance test in NumPy's operation order, angle consistency,
//      pair weight; survivors set two bits of a symmetric bit mask                   (:399-472)
//   E. builds a deterministic CSR (both directions, columns ascending) of the compact N x N affinity W
//   F. runs the fitters on per-correspondence quantities: every stacked-row sum of the reference
//      factors as sum_rows = sum_c (row-degree of c) * (per-correspondence term), the spectral
//      affinity is A = diag(h) W + W diag(h), and x = u_p u_q w_pq has degrees u .* (W u) -- so each
//      IRLS round is an O(N) reduction and each power-iteration step one CSR pass with two
//      right-hand sides and a single barrier.
// All arithmetic that decides anything (filters) or feeds the pose is float64, like the reference.
//
// Determinism: every reduction has a fixed order; atomics are used only for bit-mask OR and for slot
// allocation in the candidate list (which changes the order candidates are *processed* in, never a value
// or the position a value is stored at).  Results do not depend on which CTA/SM/GPU runs a pair.
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/rp_b200.h"

namespace {

#ifndef RP_THREADS
#define RP_THREADS 128
#endif
#ifndef RP_MIN_BLOCKS
#define RP_MIN_BLOCKS 4
#endif
constexpr int T = RP_THREADS;       // threads per CTA: small CTAs, several per SM, so the serial 4x4 eigen
constexpr int NWARP = T / 32;       // section of one pair overlaps the parallel phases of the others
constexpr int STAGE = 64;           // per-warp candidate staging entries
constexpr int KMAX = RP_MAX_TOPK;
constexpr int NSUM = 25;            // sums per Horn fit
constexpr int NUM_ALTER = 5;        // rpmodule.py:229
constexpr int NUM_REWEIGHT = 5;     // rpmodule.py:228
constexpr double OFFSET = 50.0;     // rpmodule.py:231
constexpr double EPS = 1e-12;       // rpmodule.py:232
constexpr double UNOBS_DAMP = 0.6;  // rpmodule.py:467
constexpr float FEAT_SCALING = 100.0f;  // rpmodule.py:327
// Pairs whose weight is below PRUNE_REL * (largest pair weight of the scan pair) are left out of the CSR: every sum they
// enter is dominated by terms >= 1e22 times larger, so in float64 they are exact no-ops (soft-match weights are
// row-normalised: a keypoint's best candidate has f ~ 1, its other candidates ~ e^-400).  Filters, counts and the
// surviving-pair set are unaffected; the fitters then only visit correspondences that still have a pair.
constexpr double PRUNE_REL = 1e-22;
constexpr unsigned RM_ROW = 0x3fffu, RM_CHUNK = 0x1ffu;   // PairView::rowmap fields: N <= 16383 rows, T <= 512 chunks
constexpr int RM_C0 = 14, RM_C1 = 23;
static_assert(RP_THREADS <= 512, "rowmap holds 9-bit chunk indices");

static long long g_launches = 0;

struct SolveArgs {
    int B;
    const int32_t* off_s;
    const int32_t* off_t;
    const double* pc_s; const double* nrm_s; const float* feat_s; const double* w_s;
    const double* pc_t; const double* nrm_t; const float* feat_t; const double* w_t;
    int feat_dim;
    const rp_params* params;
    const int32_t* param_idx;
    const int32_t* zero_row_topk;
    const int32_t* feat_sum_order;
    // solve-only entry (rp_spectral_irls_solve): nodes = correspondences with caller geometry, edges = caller pair list
    int solve_only;
    const double* node_wp; const double* node_wn;      // explicit base weights (NULL: row degrees of W)
    const int32_t* edge_off; const int32_t* edge_rc; const double* edge_w;
    int max_topk;
    long long edge_cap;
    char* ws;               // workspace base; first 256 bytes = header (work counter)
    size_t slot_bytes;
    size_t o_geo, o_cj, o_mask, o_edges, o_ew, o_rowstart, o_cols, o_vals;     // o_rowstart: uint16 [Nmax][NWmax] word-prefix counts
    int Nmax, NWmax;
    double* T_out; int32_t* status; int32_t* stats;
    int stop_after;
    int has_dbg;
    rp_debug dbg;
    int mask_in_smem;
    int tfeat_stride;       // odd stride (floats) of the staged target descriptors
    size_t sm_mask_off;     // byte offset of the bit mask in dynamic shared memory (when mask_in_smem)
    int dyn_in_global;      // scan pairs too large for shared memory: the per-pair vectors live in the slot (o_dyn)
    size_t o_dyn;
    int pi_switch;          // ROBUST variant: plain power steps before the accelerated iteration takes over
    int pi_fast_cap;        // fast variant: power steps per alternation after which a pair is handed to the ROBUST variant
    size_t sm_geo_off;      // 512-thread build: byte offset of the fitters' geometry copy in dynamic shared memory (0: none)
    int csr_smem_cap;       // small batches (fewer CTAs than fit an SM): CSR entries that fit the idle shared memory, else 0
    size_t sm_csr_off;
};

// ----------------------------------------------------------------------------------------------
// small device helpers
__device__ __forceinline__ double shfl_xor_d(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += shfl_xor_d(v, o);
    return v;
}

// NumPy's float32 add.reduce over a contiguous axis (pairwise_sum, 8 accumulators for 8<=n<=128;
// numpy/_core/src/umath/loops_utils.h.src) applied to (s-t)^2.  No FMA contraction anywhere.
__device__ __forceinline__ float sq_diff(const float* __restrict__ s, const float* __restrict__ t, int c) {
    float d = __fsub_rn(s[c], t[c]);
    return __fmul_rn(d, d);
}
// Same arithmetic, 128-bit shared-memory loads: requires n % 8 == 0, n <= 128 and 16-byte aligned rows.
__device__ __forceinline__ float numpy_sqdist_f32_v4(const float* __restrict__ s, const float* __restrict__ t, int n) {
    const float4* s4 = reinterpret_cast<const float4*>(s);
    const float4* t4 = reinterpret_cast<const float4*>(t);
    float r[8];
    {
        float4 a = s4[0], b = t4[0], c = s4[1], d = t4[1];
        float e;
        e = __fsub_rn(a.x, b.x); r[0] = __fmul_rn(e, e); e = __fsub_rn(a.y, b.y); r[1] = __fmul_rn(e, e);
        e = __fsub_rn(a.z, b.z); r[2] = __fmul_rn(e, e); e = __fsub_rn(a.w, b.w); r[3] = __fmul_rn(e, e);
        e = __fsub_rn(c.x, d.x); r[4] = __fmul_rn(e, e); e = __fsub_rn(c.y, d.y); r[5] = __fmul_rn(e, e);
        e = __fsub_rn(c.z, d.z); r[6] = __fmul_rn(e, e); e = __fsub_rn(c.w, d.w); r[7] = __fmul_rn(e, e);
    }
    for (int i = 2; i < (n >> 2); i += 2) {
        float4 a = s4[i], b = t4[i], c = s4[i + 1], d = t4[i + 1];
        float e;
        e = __fsub_rn(a.x, b.x); r[0] = __fadd_rn(r[0], __fmul_rn(e, e)); e = __fsub_rn(a.y, b.y); r[1] = __fadd_rn(r[1], __fmul_rn(e, e));
        e = __fsub_rn(a.z, b.z); r[2] = __fadd_rn(r[2], __fmul_rn(e, e)); e = __fsub_rn(a.w, b.w); r[3] = __fadd_rn(r[3], __fmul_rn(e, e));
        e = __fsub_rn(c.x, d.x); r[4] = __fadd_rn(r[4], __fmul_rn(e, e)); e = __fsub_rn(c.y, d.y); r[5] = __fadd_rn(r[5], __fmul_rn(e, e));
        e = __fsub_rn(c.z, d.z); r[6] = __fadd_rn(r[6], __fmul_rn(e, e)); e = __fsub_rn(c.w, d.w); r[7] = __fadd_rn(r[7], __fmul_rn(e, e));
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                     __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
}
__device__ float numpy_sqdist_f32(const float* __restrict__ s, const float* __restrict__ t, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, sq_diff(s, t, i));
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = sq_diff(s, t, j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], sq_diff(s, t, i + j));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, sq_diff(s, t, i));
    return res;
}

// When either descriptor array is not C-contiguous (the reference's own pipeline hands over transposed views:
// rpmodule.py:531-532 `torch_op.npy(interpolate(...)).T`), NumPy's nditer makes the channel axis an outer loop and the
// float32 reduction degenerates to a plain sequential sum over the channels.
__device__ float numpy_sqdist_f32_seq(const float* __restrict__ s, const float* __restrict__ t, int n) {
    float res = sq_diff(s, t, 0);
    for (int i = 1; i < n; ++i) res = __fadd_rn(res, sq_diff(s, t, i));
    return res;
}

// ---- 32-channel descriptors, one source row per thread (phase A of the fused kernel) ---------------------------------
// Packed float32 pairs (FADD2 / FMUL2 on sm_100a): subtraction and squaring are element-wise IEEE operations, so the pair
// forms give the same bits as the scalar ones with half the instructions.  The *additions* stay scalar: ptxas contracts a
// mul.rn.f32x2 feeding an add.rn.f32x2 into FFMA2 (it honours .rn only for the scalar forms), which would change dij.
__device__ __forceinline__ unsigned long long f32x2_sub(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long f32x2_sq(unsigned long long a) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(d) : "l"(a));
    return d;
}
__device__ __forceinline__ void f32x2_unpack(unsigned long long v, float& lo, float& hi) {
    lo = __uint_as_float((unsigned)(v & 0xffffffffull)); hi = __uint_as_float((unsigned)(v >> 32));
}
// sq[c] = (s[c] - t[c])^2 for the 32 channels; s in registers as 16 packed pairs, t from shared memory (16-byte aligned row)
__device__ __forceinline__ void sqdiff32(const unsigned long long (&s)[16], const float* __restrict__ t, float (&sq)[32]) {
    const ulonglong2* t2 = reinterpret_cast<const ulonglong2*>(t);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const ulonglong2 tv = t2[q];
        f32x2_unpack(f32x2_sq(f32x2_sub(s[2 * q], tv.x)), sq[4 * q], sq[4 * q + 1]);
        f32x2_unpack(f32x2_sq(f32x2_sub(s[2 * q + 1], tv.y)), sq[4 * q + 2], sq[4 * q + 3]);
    }
}
// NumPy's pairwise float32 add.reduce of 32 values (8 accumulators, see numpy_sqdist_f32) / the plain sequential sum
__device__ __forceinline__ float sum32_numpy(const float (&sq)[32]) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = sq[j];
#pragma unroll
    for (int i = 8; i < 32; i += 8)
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], sq[i + j]);
    return __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                     __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
}
__device__ __forceinline__ float sum32_seq(const float (&sq)[32]) {
    float res = sq[0];
#pragma unroll
    for (int i = 1; i < 32; ++i) res = __fadd_rn(res, sq[i]);
    return res;
}

// (a0*b0 + a1*b1) + a2*b2 with separately rounded products: NumPy's (a*b).sum(1) on [P,3].
__device__ __forceinline__ double dot3_np(double a0, double a1, double a2, double b0, double b1, double b2) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}
__device__ __forceinline__ double clip1(double x) { return fmin(fmax(x, -1.0), 1.0); }

// ----------------------------------------------------------------------------------------------
// 4x4 symmetric eigen: eigenvector of the largest eigenvalue of Horn's N (rpmodule.py:46-53).
// Cyclic Jacobi -- robust path (used when the fast path declines).
__device__ void jacobi4_max(const double A_in[4][4], double q[4]) {
    double A[4][4], V[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { A[i][j] = A_in[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            diag += A[i][i] * A[i][i];
#pragma unroll
            for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-32 * (diag + off) || off == 0.0) break;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
#pragma unroll
            for (int qq = p + 1; qq < 4; ++qq) {
                double apq = A[p][qq];
                if (apq == 0.0) continue;
                double theta = (A[qq][qq] - A[p][p]) / (2.0 * apq);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = rsqrt(t * t + 1.0);
                double s = t * c;
#pragma unroll
                for (int k = 0; k < 4; ++k) {      // A <- A J
                    double akp = A[k][p], akq = A[k][qq];
                    A[k][p] = c * akp - s * akq;
                    A[k][qq] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {      // A <- J^T A
                    double apk = A[p][k], aqk = A[qq][k];
                    A[p][k] = c * apk - s * aqk;
                    A[qq][k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    double vkp = V[k][p], vkq = V[k][qq];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][qq] = s * vkp + c * vkq;
                }
            }
        }
    }
    int best = 0;
    double bv = A[0][0];
#pragma unroll
    for (int i = 1; i < 4; ++i) if (A[i][i] > bv) { bv = A[i][i]; best = i; }
    double nn = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { q[k] = (best == 0 ? V[k][0] : best == 1 ? V[k][1] : best == 2 ? V[k][2] : V[k][3]); nn += q[k] * q[k]; }
    nn = rsqrt(nn);
#pragma unroll
    for (int k = 0; k < 4; ++k) q[k] *= nn;
}

__device__ __forceinline__ double det3(double a, double b, double c, double d, double e, double f,
                                       double g, double h, double i) {
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// Column `col` of adj(B) for symmetric 4x4 B.  adj(B) = alpha v v^T when B is singular of rank 3.
__device__ __forceinline__ void adj_col(const double B[4][4], int col, double v[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int r0 = (k == 0) ? 1 : 0, r1 = (k <= 1) ? 2 : 1, r2 = (k <= 2) ? 3 : 2;
        int c0 = (col == 0) ? 1 : 0, c1 = (col <= 1) ? 2 : 1, c2 = (col <= 2) ? 3 : 2;
        double d = det3(B[r0][c0], B[r0][c1], B[r0][c2], B[r1][c0], B[r1][c1], B[r1][c2],
                        B[r2][c0], B[r2][c1], B[r2][c2]);
        v[k] = ((k + col) & 1) ? -d : d;
    }
}

// Fast path: lambda_max by Newton on the characteristic polynomial, eigenvector from the adjugate, Rayleigh
// polish.  lam_hint > 0 warm-starts Newton at the previous fit's eigenvalue; the result is accepted only with
// a certificate that it is the LARGEST root (first and second derivative of the real-rooted quartic positive;
// the third is 24*lam) and a small eigen-residual.  Otherwise Newton restarts from the safe upper bound, and
// if that declines too (near-degenerate top eigenvalue) the caller falls back to Jacobi.
__device__ bool horn_eig_fast(const double N[4][4], double q[4], double lam_hint, double* lam_out) {
    double fro = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) fro += N[i][j] * N[i][j];
    if (!(fro > 0.0) || !isfinite(fro)) return false;
    // characteristic polynomial  l^4 + c2 l^2 + c1 l + c0  (trace N == 0)
    double c2 = -0.5 * fro;
    double m0 = det3(N[1][1], N[1][2], N[1][3], N[2][1], N[2][2], N[2][3], N[3][1], N[3][2], N[3][3]);
    double m1 = det3(N[0][0], N[0][2], N[0][3], N[2][0], N[2][2], N[2][3], N[3][0], N[3][2], N[3][3]);
    double m2 = det3(N[0][0], N[0][1], N[0][3], N[1][0], N[1][1], N[1][3], N[3][0], N[3][1], N[3][3]);
    double m3 = det3(N[0][0], N[0][1], N[0][2], N[1][0], N[1][1], N[1][2], N[2][0], N[2][1], N[2][2]);
    double c1 = -(m0 + m1 + m2 + m3);
    double k0 = det3(N[1][0], N[1][2], N[1][3], N[2][0], N[2][2], N[2][3], N[3][0], N[3][2], N[3][3]);
    double k1 = det3(N[1][0], N[1][1], N[1][3], N[2][0], N[2][1], N[2][3], N[3][0], N[3][1], N[3][3]);
    double k2 = det3(N[1][0], N[1][1], N[1][2], N[2][0], N[2][1], N[2][2], N[3][0], N[3][1], N[3][2]);
    double c0 = N[0][0] * m0 - N[0][1] * k0 + N[0][2] * k1 - N[0][3] * k2;
    const double scale = sqrt(fro);
    const double bound = 0.8660254037844387 * scale * (1.0 + 1e-12);   // >= lambda_max (traceless symmetric 4x4)
    const bool have_hint = (lam_hint > 0.0) && (lam_hint < bound);
    for (int attempt = 0; attempt < 2; ++attempt) {
        double lam = (attempt == 0 && have_hint) ? lam_hint : bound;
        bool ok = false;
        for (int it = 0; it < 60; ++it) {
            double l2 = lam * lam;
            double p = (l2 + c2) * l2 + c1 * lam + c0;
            double dp = (4.0 * l2 + 2.0 * c2) * lam + c1;
            if (!(dp > 0.0)) break;
            double step = p / dp;
            lam -= step;
            if (fabs(step) <= 4.0e-16 * fabs(lam)) { ok = true; break; }
            if (!(lam > 0.0) || !(lam <= 2.0 * bound)) break;
        }
        if (ok) {   // largest-root certificate
            double dp = (4.0 * lam * lam + 2.0 * c2) * lam + c1;
            double ddp = 12.0 * lam * lam + 2.0 * c2;
            ok = (dp > 0.0) && (ddp > 0.0) && (lam > 0.0);
        }
        if (ok) {
            double v[4];
            bool good = true;
            for (int polish = 0; polish < 2 && good; ++polish) {
                double Bm[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) Bm[i][j] = N[i][j] - (i == j ? lam : 0.0);
                // diagonal of adj(B) = alpha v_i^2: take the column with the largest diagonal cofactor
                double d0 = det3(Bm[1][1], Bm[1][2], Bm[1][3], Bm[2][1], Bm[2][2], Bm[2][3], Bm[3][1], Bm[3][2], Bm[3][3]);
                double d1 = det3(Bm[0][0], Bm[0][2], Bm[0][3], Bm[2][0], Bm[2][2], Bm[2][3], Bm[3][0], Bm[3][2], Bm[3][3]);
                double d2 = det3(Bm[0][0], Bm[0][1], Bm[0][3], Bm[1][0], Bm[1][1], Bm[1][3], Bm[3][0], Bm[3][1], Bm[3][3]);
                double d3 = det3(Bm[0][0], Bm[0][1], Bm[0][2], Bm[1][0], Bm[1][1], Bm[1][2], Bm[2][0], Bm[2][1], Bm[2][2]);
                int col = 0; double dm = fabs(d0);
                if (fabs(d1) > dm) { dm = fabs(d1); col = 1; }
                if (fabs(d2) > dm) { dm = fabs(d2); col = 2; }
                if (fabs(d3) > dm) { dm = fabs(d3); col = 3; }
                // |adj| ~ prod(lambda_max - lambda_i): tiny => (near-)degenerate top eigenvalue
                if (!(dm > 1e-9 * scale * scale * scale)) { good = false; break; }
                if (col == 0) adj_col(Bm, 0, v); else if (col == 1) adj_col(Bm, 1, v);
                else if (col == 2) adj_col(Bm, 2, v); else adj_col(Bm, 3, v);
                double nn = rsqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] *= nn;
                double rq = 0.0;                 // Rayleigh quotient refines lambda (cubic convergence)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double sx = 0.0;
#pragma unroll
                    for (int j = 0; j < 4; ++j) sx += N[i][j] * v[j];
                    rq += v[i] * sx;
                }
                lam = rq;
                if (polish == 0) {               // usually already converged: skip the second adjugate
                    double res0 = 0.0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        double sx = -lam * v[i];
#pragma unroll
                        for (int j = 0; j < 4; ++j) sx += N[i][j] * v[j];
                        res0 += sx * sx;
                    }
                    if (res0 <= 1e-29 * fro) break;
                }
            }
            if (good) {
                double res = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    double sx = -lam * v[i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) sx += N[i][j] * v[j];
                    res += sx * sx;
                }
                if (res <= 1e-22 * fro) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) q[k] = v[k];
                    *lam_out = lam;
                    return true;
                }
            }
        }
        if (!have_hint) break;   // the bound start was the first attempt
    }
    return false;
}

__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));     // one MUFU.RSQ, ~2 ulp: far inside the pre-test margin
    return r;
}

// ----------------------------------------------------------------------------------------------
struct Pose { double R[9]; double t[3]; double sm[3]; double tm[3]; };

struct Shared {
    double red[2][NWARP][NSUM];
    double fin[32];
    Pose pose;
    double scal[8];      // 0: lambda warm start, 1: prefilter margin, 2: CSR weight threshold
    unsigned long long wmax_bits;   // largest pair weight (non-negative doubles order like their bit patterns)
    int cnt[8];          // 0 candidate count, 1 M1, 2 M2, 3 scan carry, 4 pair id, 5 nnz
    int ties;            // source rows whose top-k SET is not determined by the keys (see stats[6])
    int warp_tot[NWARP];
    int warp_tot2[NWARP];
    int crow0[T];        // merge-path: compact (non-empty) row index of the first non-zero of thread t's chunk
    double part[T][4];   // merge-path partial sums: [0..1] leading piece of a row begun in an earlier chunk,
                         //                          [2..3] trailing piece of a row that continues in later chunks
    unsigned stage[NWARP][STAGE];
    __align__(16) float sfeat[NWARP][RP_MAX_FEAT_DIM];
};

// Per-correspondence geometry, SoA in the slot's global workspace (stride = Nmax doubles).
enum { G_PX = 0, G_PY, G_PZ, G_QX, G_QY, G_QZ, G_NX, G_NY, G_NZ, G_MX, G_MY, G_MZ, G_WS, G_WT, G_F, G_DEG,
       G_T0, G_T1,        // scratch N-vectors of the accelerated eigen iteration (search direction p and A p)
       G_COUNT };

struct PairView {
    int ns, nt, K, N, NW;
    double* geo; int gstride;
    const double* hgeo; int hgs;   // the 12 position / normal vectors the fitters read (G_PX..G_MZ): pv.geo, or a shared-memory copy
    int* cj;
    unsigned* mask;
    unsigned* edges; double* ew;
    int* rowstart;       // [N+1] shared memory
    uint16_t* cols; double* vals;   // lane-interleaved: logical entry k lives at (k % E) * T + k / E
    uint16_t* cols_s; double* vals_s; int ism;   // shared-memory copy of the entries with k % E < ism (same indexing)
                         // cols word: column (14 bits) | bit 15 = first entry of its row | bit 14 = last entry
    int E, nnz, nrows;   // non-zeros per thread chunk, total directed non-zeros, rows with non-zeros
    unsigned* rowmap;    // [nrows] shared: row id (14 bits) | first chunk << 14 (9 bits) | last chunk << 23 (9 bits)
    double* S;           // [2*nrows] shared: row sums (s1,s2) of rows lying inside one chunk
    // shared-memory vectors
    double *aP, *aN, *res, *ua, *ub, *sv;
    float4* sp4;         // [ns] float32 source keypoint positions   (phase C only; aliases the vectors)
    float4* tq4;         // [N]  float32 target position of each correspondence
};

// One weighted Horn fit from per-correspondence weights aP (positions, incl. mu) / aN (normals):
// centroids (rpmodule.py:240-243), 3x3 cross-covariance (:245-249,39-43), Horn N (:46-49), eigen, R (:54-56),
// t (:250).  `mean_div`: centroid weights are aP/mean_div (fit_horn87 / fit_spectral's first fit use the
// un-scaled pair weights for the centroids, rpmodule.py:72-75,107-110).
__device__ void horn_fit(Shared& sh, const PairView& pv, double mean_div, int& red_buf) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double* geo = pv.hgeo; const int gs = pv.hgs;
    if (NWARP == 4) {
        // warp-specialised: each warp owns 6-7 of the 25 sums over ALL correspondences, so the only
        // cross-lane traffic is one butterfly per owned sum and there is no cross-warp combine.
        double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0;
        if (warp == 0) {
            for (int cr = lane; cr < pv.nrows; cr += 32) {
                const int c = pv.rowmap[cr] & RM_ROW;
                double wp = pv.aP[c];
                a0 += wp;
                a1 += wp * geo[G_PX * gs + c]; a2 += wp * geo[G_PY * gs + c]; a3 += wp * geo[G_PZ * gs + c];
                a4 += wp * geo[G_QX * gs + c]; a5 += wp * geo[G_QY * gs + c]; a6 += wp * geo[G_QZ * gs + c];
            }
        } else if (warp == 1) {
            for (int cr = lane; cr < pv.nrows; cr += 32) {
                const int c = pv.rowmap[cr] & RM_ROW;
                double wp = pv.aP[c];
                double qx = geo[G_QX * gs + c], qy = geo[G_QY * gs + c], qz = geo[G_QZ * gs + c];
                double wpx = wp * geo[G_PX * gs + c], wpy = wp * geo[G_PY * gs + c];
                a0 += wpx * qx; a1 += wpx * qy; a2 += wpx * qz; a3 += wpy * qx; a4 += wpy * qy; a5 += wpy * qz;
            }
        } else if (warp == 2) {
            for (int cr = lane; cr < pv.nrows; cr += 32) {
                const int c = pv.rowmap[cr] & RM_ROW;
                double wpz = pv.aP[c] * geo[G_PZ * gs + c];
                double wnx = pv.aN[c] * geo[G_NX * gs + c];
                double mx = geo[G_MX * gs + c], my = geo[G_MY * gs + c], mz = geo[G_MZ * gs + c];
                a0 += wpz * geo[G_QX * gs + c]; a1 += wpz * geo[G_QY * gs + c]; a2 += wpz * geo[G_QZ * gs + c];
                a3 += wnx * mx; a4 += wnx * my; a5 += wnx * mz;
            }
        } else {
            for (int cr = lane; cr < pv.nrows; cr += 32) {
                const int c = pv.rowmap[cr] & RM_ROW;
                double wn = pv.aN[c];
                double wny = wn * geo[G_NY * gs + c], wnz = wn * geo[G_NZ * gs + c];
                double mx = geo[G_MX * gs + c], my = geo[G_MY * gs + c], mz = geo[G_MZ * gs + c];
                a0 += wny * mx; a1 += wny * my; a2 += wny * mz; a3 += wnz * mx; a4 += wnz * my; a5 += wnz * mz;
            }
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4); a5 = warp_sum(a5);
        if (warp == 0) a6 = warp_sum(a6);
        if (lane == 0) {
            // fin layout: 0 W | 1-3 sum wp p | 4-6 sum wp q | 7-15 sum wp p q^T | 16-24 sum wn n m^T
            const int base = (warp == 0) ? 0 : (warp == 1) ? 7 : (warp == 2) ? 13 : 19;
            sh.fin[base + 0] = a0; sh.fin[base + 1] = a1; sh.fin[base + 2] = a2;
            sh.fin[base + 3] = a3; sh.fin[base + 4] = a4; sh.fin[base + 5] = a5;
            if (warp == 0) sh.fin[6] = a6;
        }
        __syncthreads();
    } else if (NWARP >= 13) {
        // 512-thread build: every one of the 25 sums is sum_c w[c] X[c] Y[c] with w in {aP, aN}, X in {1, p_a, n_a}, Y in {1, q_b,
        // m_b}; warp k takes sums k and k + NWARP over ALL correspondences (operands in shared memory), so a fit costs two short
        // loops and two butterflies per warp instead of 25 butterflies per warp and a cross-warp combine.
        static_assert(NWARP < 13 || (NWARP <= 16 && 2 * NWARP >= NSUM), "sum k < NWARP weighs with aP, its partner k + NWARP >= 16 with aN");
        auto rows_of = [&](int k, const double*& xv, const double*& yv, bool& hx, bool& hy) {
            int xa = -1, ya = -1;
            if (k >= 1 && k <= 3) xa = G_PX + (k - 1);
            else if (k >= 4 && k <= 6) ya = G_QX + (k - 4);
            else if (k >= 7 && k <= 15) { xa = G_PX + (k - 7) / 3; ya = G_QX + (k - 7) % 3; }
            else if (k >= 16) { xa = G_NX + (k - 16) / 3; ya = G_MX + (k - 16) % 3; }
            hx = xa >= 0; hy = ya >= 0;
            xv = geo + (size_t)(hx ? xa : 0) * gs; yv = geo + (size_t)(hy ? ya : 0) * gs;
        };
        const int k0 = warp, k1 = warp + NWARP;                // both sums of the warp in ONE loop: independent loads overlap
        const bool two = k1 < NSUM;
        const double *x0, *y0, *x1, *y1; bool hx0, hy0, hx1, hy1;
        rows_of(k0, x0, y0, hx0, hy0);
        rows_of(two ? k1 : 16, x1, y1, hx1, hy1);
        const double* w0 = k0 < 16 ? pv.aP : pv.aN;
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 4
        for (int cr = lane; cr < pv.nrows; cr += 32) {
            const int c = pv.rowmap[cr] & RM_ROW;
            double t0 = w0[c];
            if (hx0) t0 *= x0[c];
            if (hy0) t0 *= y0[c];
            acc0 += t0;
            if (two) acc1 += (pv.aN[c] * x1[c]) * y1[c];
        }
        acc0 = warp_sum(acc0);
        if (two) acc1 = warp_sum(acc1);
        if (lane == 0) { sh.fin[k0] = acc0; if (two) sh.fin[k1] = acc1; }
        __syncthreads();
    } else {
    double acc[NSUM];
#pragma unroll
    for (int k = 0; k < NSUM; ++k) acc[k] = 0.0;
    for (int cr = tid; cr < pv.nrows; cr += T) {
        const int c = pv.rowmap[cr] & RM_ROW;
        double wp = pv.aP[c], wn = pv.aN[c];
        double px = geo[G_PX * gs + c], py = geo[G_PY * gs + c], pz = geo[G_PZ * gs + c];
        double qx = geo[G_QX * gs + c], qy = geo[G_QY * gs + c], qz = geo[G_QZ * gs + c];
        double nx = geo[G_NX * gs + c], ny = geo[G_NY * gs + c], nz = geo[G_NZ * gs + c];
        double mx = geo[G_MX * gs + c], my = geo[G_MY * gs + c], mz = geo[G_MZ * gs + c];
        acc[0] += wp;
        double wpx = wp * px, wpy = wp * py, wpz = wp * pz;
        acc[1] += wpx; acc[2] += wpy; acc[3] += wpz;
        acc[4] += wp * qx; acc[5] += wp * qy; acc[6] += wp * qz;
        acc[7] += wpx * qx; acc[8] += wpx * qy; acc[9] += wpx * qz;
        acc[10] += wpy * qx; acc[11] += wpy * qy; acc[12] += wpy * qz;
        acc[13] += wpz * qx; acc[14] += wpz * qy; acc[15] += wpz * qz;
        double wnx = wn * nx, wny = wn * ny, wnz = wn * nz;
        acc[16] += wnx * mx; acc[17] += wnx * my; acc[18] += wnx * mz;
        acc[19] += wny * mx; acc[20] += wny * my; acc[21] += wny * mz;
        acc[22] += wnz * mx; acc[23] += wnz * my; acc[24] += wnz * mz;
    }
#pragma unroll
    for (int k = 0; k < NSUM; ++k) acc[k] = warp_sum(acc[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NSUM; ++k) sh.red[red_buf][warp][k] = acc[k];
    }
    __syncthreads();
    if (warp == 0 && lane < NSUM) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) s += sh.red[red_buf][w][lane];
        sh.fin[lane] = s;
    }
    __syncthreads();
    }
    if (warp == 0) {
        if (lane == 0) {
            const double* f = sh.fin;
            double W = f[0];
            double Wm = W / mean_div;
            double inv = 1.0 / (Wm + EPS);                       // centroids  sum(w p) / (sum(w) + EPS)
            double sm[3], tm[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) { sm[a] = (f[1 + a] / mean_div) * inv; tm[a] = (f[4 + a] / mean_div) * inv; }
            // M_ab = sum wp (p_a - sm_a)(q_b - tm_b) + sum wn n_a m_b
            double M[3][3];
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    M[a][b] = f[7 + 3 * a + b] - sm[a] * f[4 + b] - f[1 + a] * tm[b] + W * sm[a] * tm[b] + f[16 + 3 * a + b];
            double Nm[4][4];
            Nm[0][0] = M[0][0] + M[1][1] + M[2][2];
            Nm[0][1] = M[1][2] - M[2][1]; Nm[0][2] = M[2][0] - M[0][2]; Nm[0][3] = M[0][1] - M[1][0];
            Nm[1][1] = M[0][0] - M[1][1] - M[2][2];
            Nm[1][2] = M[0][1] + M[1][0]; Nm[1][3] = M[0][2] + M[2][0];
            Nm[2][2] = M[1][1] - M[0][0] - M[2][2];
            Nm[2][3] = M[1][2] + M[2][1];
            Nm[3][3] = M[2][2] - M[0][0] - M[1][1];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < i; ++j) Nm[i][j] = Nm[j][i];
            double q[4];
            double lam_new = 0.0;
            if (!horn_eig_fast(Nm, q, sh.scal[0], &lam_new)) { jacobi4_max(Nm, q); lam_new = 0.0; }
            sh.scal[0] = lam_new;                                // warm start for this pair's next fit
            double a = q[0], b = q[1], c = q[2], d = q[3];
            Pose& P = sh.pose;
            P.R[0] = a * a + b * b - c * c - d * d; P.R[1] = 2 * (b * c - a * d); P.R[2] = 2 * (b * d + a * c);
            P.R[3] = 2 * (c * b + a * d); P.R[4] = a * a - b * b + c * c - d * d; P.R[5] = 2 * (c * d - a * b);
            P.R[6] = 2 * (d * b - a * c); P.R[7] = 2 * (d * c + a * b); P.R[8] = a * a - b * b - c * c + d * d;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                P.t[r] = -(P.R[3 * r] * sm[0] + P.R[3 * r + 1] * sm[1] + P.R[3 * r + 2] * sm[2]) + tm[r];
                P.sm[r] = sm[r]; P.tm[r] = tm[r];
            }
        }
    }
    __syncthreads();
    red_buf ^= 1;
}

// Per-correspondence residuals w.r.t. the current pose (rpmodule.py:252-253); optionally reweight (:255).
// res[c] = mu*||R(p-sm)-(q-tm)||^2 + ||R n - m||^2  (the r of :262-263).
__device__ void residual_pass(const Shared& sh, const PairView& pv, double mu, bool reweight) {
    const Pose& P = sh.pose;
    const double* geo = pv.hgeo; const int gs = pv.hgs;
    for (int cr = threadIdx.x; cr < pv.nrows; cr += T) {
        const int c = pv.rowmap[cr] & RM_ROW;       // only correspondences that still have a pair carry weight
        double px = geo[G_PX * gs + c] - P.sm[0], py = geo[G_PY * gs + c] - P.sm[1], pz = geo[G_PZ * gs + c] - P.sm[2];
        double qx = geo[G_QX * gs + c] - P.tm[0], qy = geo[G_QY * gs + c] - P.tm[1], qz = geo[G_QZ * gs + c] - P.tm[2];
        double dx = P.R[0] * px + P.R[1] * py + P.R[2] * pz - qx;
        double dy = P.R[3] * px + P.R[4] * py + P.R[5] * pz - qy;
        double dz = P.R[6] * px + P.R[7] * py + P.R[8] * pz - qz;
        double rp = mu * (dx * dx + dy * dy + dz * dz);
        double nx = geo[G_NX * gs + c], ny = geo[G_NY * gs + c], nz = geo[G_NZ * gs + c];
        double ex = P.R[0] * nx + P.R[1] * ny + P.R[2] * nz - geo[G_MX * gs + c];
        double ey = P.R[3] * nx + P.R[4] * ny + P.R[5] * nz - geo[G_MY * gs + c];
        double ez = P.R[6] * nx + P.R[7] * ny + P.R[8] * nz - geo[G_MZ * gs + c];
        double rn = ex * ex + ey * ey + ez * ez;
        pv.res[c] = rp + rn;
        if (reweight) {
            // w/(1+r) as w * (1/(1+r)): the weights reach the denormal range (e^-400-ish products), which would send
            // every division down the slow path; 1+r is always a normal number >= 1.  (<= 1 ulp from the quotient.)
            pv.aP[c] = pv.aP[c] * (1.0 / (1.0 + rp));
            pv.aN[c] = pv.aN[c] * (1.0 / (1.0 + rn));
        }
    }
    __syncthreads();
}

__device__ void irls_rounds(Shared& sh, const PairView& pv, double mu, int& red_buf) {
    for (int it = 0; it < NUM_REWEIGHT; ++it) {
        horn_fit(sh, pv, 1.0, red_buf);
        residual_pass(sh, pv, mu, true);
    }
}

// Block reduction of NV values, result broadcast to every thread (one __syncthreads, double buffered).
template <int NV>
__device__ void block_sum(Shared& sh, double (&v)[NV], int& red_buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) sh.red[red_buf][warp][k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NWARP; ++w) s += sh.red[red_buf][w][k];
        v[k] = s;
    }
    red_buf ^= 1;
}

// Merge-path walk over the CSR of W: every thread owns E consecutive logical non-zeros (perfect balance no
// matter how skewed the row lengths are -- inlier correspondences have ~10x the degree of outliers).
// Row boundaries are flag bits in the column word (reset / store are selects and predicated stores).  F::term(c, w, s1, s2) accumulates one non-zero;
// after a barrier F::fin(row, s1, s2) runs once per non-empty row, one thread per row.  A row split across
// chunks is summed trailing piece + leading pieces in chunk order, so the summation order is fixed.
// Rows without non-zeros are never visited.  Contains one __syncthreads.
constexpr int CSR_PB = 8;           // non-zeros fetched per batch by csr_walk

template <class F>
__device__ __forceinline__ void csr_walk(Shared& sh, const PairView& pv, F& f) {
    const int t = threadIdx.x;
    const int E = pv.E;
    int n = pv.nnz - t * E;
    n = n < 0 ? 0 : (n > E ? E : n);                      // non-zeros in this thread's chunk
    if (n > 0) {
        double* __restrict__ S = pv.S;
        double* mypart = sh.part[t];
        int cr = sh.crow0[t];
        bool headless = true;                             // current row began in an earlier chunk
        double s1 = 0.0, s2 = 0.0;
        unsigned cw = 0;
        // The non-zeros are fetched in batches of CSR_PB independent loads before any of them is used (the loop-carried
        // row state otherwise serialises one memory round trip per non-zero); the leading pv.ism entries of every chunk
        // come from shared memory when the launch had room for them (make_launch_plan), the rest from the slot.
        for (int i0 = 0; i0 < n; i0 += CSR_PB) {
            const bool from_s = i0 + CSR_PB <= pv.ism;
            const uint16_t* __restrict__ cp = (from_s ? pv.cols_s : pv.cols) + t + (size_t)i0 * T;
            const double* __restrict__ vp = (from_s ? pv.vals_s : pv.vals) + t + (size_t)i0 * T;
            unsigned cwb[CSR_PB]; double wb[CSR_PB];
#pragma unroll
            for (int u = 0; u < CSR_PB; ++u) {
                const bool ok = i0 + u < n;
                cwb[u] = ok ? (unsigned)cp[(size_t)u * T] : 0u;
                wb[u] = ok ? vp[(size_t)u * T] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < CSR_PB; ++u) {
                if (i0 + u < n) {
                    cw = cwb[u];
                    const double w = wb[u];
                    const bool st = (cw & 0x8000u) != 0, en = (cw & 0x4000u) != 0;
                    s1 = st ? 0.0 : s1; s2 = st ? 0.0 : s2;
                    cr += (st && (i0 + u) > 0) ? 1 : 0;
                    headless = headless && !st;
                    f.term((int)(cw & 0x3fffu), w, s1, s2);
                    if (en) { double* d = headless ? mypart : (S + 2 * cr); d[0] = s1; d[1] = s2; }
                }
            }
        }
        if (!(cw & 0x4000u)) { double* d = headless ? mypart : (mypart + 2); d[0] = s1; d[1] = s2; }
    }
    __syncthreads();
    for (int cr = t; cr < pv.nrows; cr += T) {
        const unsigned info = pv.rowmap[cr];
        const int p = info & RM_ROW, t0 = (info >> RM_C0) & RM_CHUNK, t1 = info >> RM_C1;
        double s1, s2;
        if (t0 == t1) { s1 = pv.S[2 * cr]; s2 = pv.S[2 * cr + 1]; }
        else {
            s1 = sh.part[t0][2]; s2 = sh.part[t0][3];
            for (int tt = t0 + 1; tt <= t1; ++tt) { s1 += sh.part[tt][0]; s2 += sh.part[tt][1]; }
        }
        f.fin(p, s1, s2);
    }
}

template <bool USE_SV>
struct PowerStep {            // y = S (diag(h) W + W diag(h)) S u ; accumulates ||y||^2 and ||y - lam_prev u||^2
    const double* __restrict__ cur; double* __restrict__ nxt;
    const double* __restrict__ h; const double* __restrict__ sv;
    double inv, lam_prev, acc0, acc1;
    __device__ __forceinline__ void term(int c, double w, double& s1, double& s2) const {
        double gc = cur[c];
        if (USE_SV) gc *= sv[c];
        s1 += w * gc; s2 += w * (h[c] * gc);
    }
    __device__ __forceinline__ void fin(int p, double s1, double s2) {
        double y = (h[p] * s1 + s2) * inv;           // (A u_k)_p
        if (USE_SV) y *= sv[p];
        nxt[p] = y;
        double r = y - lam_prev * (cur[p] * inv);    // residual against the previous eigenvalue estimate
        acc0 += y * y; acc1 += r * r;
    }
};

struct XDegStep {             // xdeg_p = u_p (W u)_p
    const double* __restrict__ u; double* __restrict__ aP; double* __restrict__ aN; double mu;
    __device__ __forceinline__ void term(int c, double w, double& s1, double&) const { s1 += w * u[c]; }
    __device__ __forceinline__ void fin(int p, double s1, double) const { double xd = u[p] * s1; aP[p] = mu * xd; aN[p] = xd; }
};

struct DegStep {              // row sums of W
    double* __restrict__ deg;
    __device__ __forceinline__ void term(int, double w, double& s1, double&) const { s1 += w; }
    __device__ __forceinline__ void fin(int p, double s1, double) const { deg[p] = s1; }
};

template <bool USE_SV>
struct MatVecStep {           // out = S (diag(h) W + W diag(h)) S in
    const double* __restrict__ in; double* __restrict__ out;
    const double* __restrict__ h; const double* __restrict__ sv;
    __device__ __forceinline__ void term(int c, double w, double& s1, double& s2) const {
        double gc = in[c];
        if (USE_SV) gc *= sv[c];
        s1 += w * gc; s2 += w * (h[c] * gc);
    }
    __device__ __forceinline__ void fin(int p, double s1, double s2) {
        double y = h[p] * s1 + s2;
        if (USE_SV) y *= sv[p];
        out[p] = y;
    }
};

// Largest eigenpair of the symmetric Rayleigh-Ritz matrix of the accelerated eigen iteration (3x3, or its leading 2x2 block
// when n == 2).  2x2: closed form.  3x3: Newton on the characteristic cubic from the Gershgorin upper bound (for a real-rooted
// polynomial the iterates decrease monotonically to the largest root), eigenvector = the largest cross product of two rows of
// B - theta I, one Rayleigh-quotient refinement.  ~200 flops with short dependency chains instead of the ~1500 of a Jacobi
// solve; returns false (caller falls back to Jacobi) when Newton does not settle or the eigenvector is not resolved.
__device__ __forceinline__ bool ritz_max(const double B[4][4], int n, double a[3], double* theta_out) {
    if (n == 2) {
        const double p = B[0][0], q = B[1][1], o = B[0][1];
        const double hd = 0.5 * (p - q), rad = sqrt(hd * hd + o * o);
        const double th = 0.5 * (p + q) + rad;
        // eigenvector (o, th - p) or (th - q, o): take the one built from the larger difference
        double v0, v1;
        if (hd >= 0.0) { v0 = th - q; v1 = o; } else { v0 = o; v1 = th - p; }
        const double n2 = v0 * v0 + v1 * v1;
        if (!(n2 > 0.0) || !isfinite(n2)) return false;
        const double inv = rsqrt(n2);
        a[0] = v0 * inv; a[1] = v1 * inv; a[2] = 0.0;
        *theta_out = th;
        return true;
    }
    const double b00 = B[0][0], b11 = B[1][1], b22 = B[2][2], b01 = B[0][1], b02 = B[0][2], b12 = B[1][2];
    const double ub = fmax(fmax(b00 + fabs(b01) + fabs(b02), b11 + fabs(b01) + fabs(b12)), b22 + fabs(b02) + fabs(b12));
    const double sc = fabs(b00) + fabs(b11) + fabs(b22) + fabs(b01) + fabs(b02) + fabs(b12);
    if (!(sc > 0.0) || !isfinite(sc)) return false;
    // shifted matrix C = B - ub I (eigenvalues <= 0): the cubic is evaluated in the small quantity mu = theta - ub
    const double c00 = b00 - ub, c11 = b11 - ub, c22 = b22 - ub;
    const double k2 = c00 + c11 + c22;                                                          // trace
    const double k1 = (c00 * c11 - b01 * b01) + (c00 * c22 - b02 * b02) + (c11 * c22 - b12 * b12);
    const double k0 = c00 * (c11 * c22 - b12 * b12) - b01 * (b01 * c22 - b12 * b02) + b02 * (b01 * b12 - c11 * b02);   // det
    double mu = 0.0;                                       // p(mu) = mu^3 - k2 mu^2 + k1 mu - k0, largest root <= 0
    bool ok = false;
    for (int it = 0; it < 80; ++it) {
        const double pm = ((mu - k2) * mu + k1) * mu - k0;
        const double dp = (3.0 * mu - 2.0 * k2) * mu + k1;
        if (!(dp > 0.0)) { ok = (pm == 0.0); break; }
        const double step = pm / dp;
        mu -= step;
        if (fabs(step) <= 2.0e-16 * sc || (it > 0 && step <= 0.0)) { ok = true; break; }   // (a negative step is rounding noise)
    }
    if (!ok) return false;
    double th = ub + mu;
    double v[3];
    for (int polish = 0; polish < 2; ++polish) {
        const double r0[3] = {b00 - th, b01, b02}, r1[3] = {b01, b11 - th, b12}, r2[3] = {b02, b12, b22 - th};
        const double x0[3] = {r0[1] * r1[2] - r0[2] * r1[1], r0[2] * r1[0] - r0[0] * r1[2], r0[0] * r1[1] - r0[1] * r1[0]};
        const double x1[3] = {r0[1] * r2[2] - r0[2] * r2[1], r0[2] * r2[0] - r0[0] * r2[2], r0[0] * r2[1] - r0[1] * r2[0]};
        const double x2[3] = {r1[1] * r2[2] - r1[2] * r2[1], r1[2] * r2[0] - r1[0] * r2[2], r1[0] * r2[1] - r1[1] * r2[0]};
        const double n0 = x0[0] * x0[0] + x0[1] * x0[1] + x0[2] * x0[2];
        const double n1 = x1[0] * x1[0] + x1[1] * x1[1] + x1[2] * x1[2];
        const double n2 = x2[0] * x2[0] + x2[1] * x2[1] + x2[2] * x2[2];
        double nb = n0; v[0] = x0[0]; v[1] = x0[1]; v[2] = x0[2];
        if (n1 > nb) { nb = n1; v[0] = x1[0]; v[1] = x1[1]; v[2] = x1[2]; }
        if (n2 > nb) { nb = n2; v[0] = x2[0]; v[1] = x2[1]; v[2] = x2[2]; }
        // |cross| ~ (theta - theta_2)(theta - theta_3): unresolved when the two leading Ritz values coincide to rounding
        if (!(nb > 1e-24 * sc * sc * sc * sc) || !isfinite(nb)) return false;
        const double inv = rsqrt(nb);
        v[0] *= inv; v[1] *= inv; v[2] *= inv;
        const double w0 = b00 * v[0] + b01 * v[1] + b02 * v[2], w1 = b01 * v[0] + b11 * v[1] + b12 * v[2], w2 = b02 * v[0] + b12 * v[1] + b22 * v[2];
        th = v[0] * w0 + v[1] * w1 + v[2] * w2;            // Rayleigh quotient: second-order accurate in the vector error
    }
    a[0] = v[0]; a[1] = v[1]; a[2] = v[2];
    *theta_out = th;
    return true;
}

constexpr int PI_SWITCH = 16;       // ROBUST variant: plain power steps before the accelerated iteration takes over
constexpr int PI_FAST_CAP = 48;     // fast variant: power steps per alternation after which a pair is handed to the ROBUST variant (which
                                    // redoes the whole pair).  Measured on B200 (32 dense 9 k-edge pairs / the 4 096-pair headline batch,
                                    // whose pairs need <= 13 steps): cap 192, switch 48 -> 6.9 ms / 720 k pairs/s; 96, 48 -> 5.3 ms; 48, 48
                                    // -> 3.5 ms; 48, 16 -> 2.8 ms / 719 k; 32, 16 -> 2.6 ms.  A pair that needs ~70 steps per alternation
                                    // costs the same either way (the redo is worth ~60 steps); beyond that the accelerated iteration wins.
                                    // RP_PI_FAST_CAP / RP_PI_SWITCH in the environment override both (tuning aid).
constexpr int RP_STATUS_RETRY = -100;   // internal: pair waits for the ROBUST pass (never visible to the caller)

// Continuation of the power iteration for graphs whose two leading eigenvalues nearly coincide (two weakly coupled
// groups of mutually consistent correspondences: lambda_2/lambda_1 -> 1, where the power method needs ~1/(1 - ratio)
// steps and ARPACK's restarted Arnoldi, which the reference calls (rpmodule.py:134,273), does not care).  Locally
// optimal CG: every step takes the best Rayleigh-Ritz vector in span{x, r, p} -- the iterate, its (orthonormalised)
// residual and the previous update -- so the error contracts like 1 - 2 sqrt(gap) instead of 1 - gap; one CSR pass
// (A r) per step, A x and A p follow by linearity.  The 3x3 projected problem is expressed in the orthonormal basis
// {x, r, p_perp} through inner products only (one 8-value reduction) and solved redundantly by every thread with the
// Jacobi routine of the Horn fit.  On entry pv.ua holds the unit iterate of the power phase; on return the unit
// eigenvector.  A "converged" residual that was reached through the A x recurrence is re-checked with a true product.
// (Only instantiated in the ROBUST variant of the kernel, see rp_solve_kernel: inlined into the fused kernel it costs the
//  common path ~9 % through register spills, as a call it forces the per-pair view into local memory.)
template <bool USE_SV>
__device__ __forceinline__ int lopcg_continue(Shared& sh, const PairView& pv, double tol, int max_it, int& red_buf, int* converged) {
    const int tid = threadIdx.x;
    const int N = pv.N;
    double* __restrict__ x = pv.ua; double* __restrict__ r = pv.ub;
    double* __restrict__ Ar = pv.aP; double* __restrict__ Ax = pv.aN;
    double* __restrict__ p = pv.geo + (size_t)G_T0 * pv.gstride;
    double* __restrict__ Ap = pv.geo + (size_t)G_T1 * pv.gstride;
    for (int c = tid; c < N; c += T) { r[c] = 0.0; Ar[c] = 0.0; Ax[c] = 0.0; p[c] = 0.0; Ap[c] = 0.0; }
    __syncthreads();
    MatVecStep<USE_SV> mv; mv.h = pv.res; mv.sv = pv.sv;
    double lam = 0.0;
    auto true_product = [&]() {                          // Ax = A x, lam = x . Ax
        mv.in = x; mv.out = Ax;
        csr_walk(sh, pv, mv);
        __syncthreads();
        double v[1] = {0.0};
        for (int c = tid; c < N; c += T) v[0] += x[c] * Ax[c];
        block_sum<1>(sh, v, red_buf);
        lam = v[0];
    };
    true_product();
    bool have_p = false, fresh = true;
    int it = 1;
    *converged = 0;
    while (it < max_it) {
        double v2[2] = {0.0, 0.0};
        for (int c = tid; c < N; c += T) { const double rc = Ax[c] - lam * x[c]; r[c] = rc; v2[0] += rc * rc; v2[1] += x[c] * rc; }
        block_sum<2>(sh, v2, red_buf);
        const double rn2 = v2[0] - v2[1] * v2[1];
        if (v2[0] <= tol * tol * lam * lam || !(rn2 > 0.0)) {
            if (fresh) { *converged = 1; break; }
            true_product(); ++it;                        // re-check against a true product, restart the direction
            fresh = true; have_p = false;
            continue;
        }
        const double rinv = rsqrt(rn2);
        for (int c = tid; c < N; c += T) r[c] = (r[c] - v2[1] * x[c]) * rinv;
        __syncthreads();                                 // r complete before it is gathered
        mv.in = r; mv.out = Ar;
        csr_walk(sh, pv, mv);
        __syncthreads();
        ++it;
        double d[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};     // xAr rAr | xAp rAp pAp xp rp pp
        for (int c = tid; c < N; c += T) {
            const double xc = x[c], rc = r[c], arc = Ar[c];
            d[0] += xc * arc; d[1] += rc * arc;
            if (have_p) {
                const double pc = p[c], apc = Ap[c];
                d[2] += xc * apc; d[3] += rc * apc; d[4] += pc * apc; d[5] += xc * pc; d[6] += rc * pc; d[7] += pc * pc;
            }
        }
        block_sum<8>(sh, d, red_buf);
        double B[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) B[i][j] = 0.0;
        B[0][0] = lam; B[0][1] = B[1][0] = d[0]; B[1][1] = d[1];
        bool use_p = false;
        double ninv = 0.0;
        if (have_p) {
            const double n2 = d[7] - d[5] * d[5] - d[6] * d[6];
            if (d[7] > 0.0 && n2 > 1e-20 * d[7]) {
                use_p = true;
                ninv = rsqrt(n2);
                B[0][2] = B[2][0] = (d[2] - d[5] * lam - d[6] * d[0]) * ninv;
                B[1][2] = B[2][1] = (d[3] - d[5] * d[0] - d[6] * d[1]) * ninv;
                B[2][2] = (d[4] - 2.0 * d[5] * d[2] - 2.0 * d[6] * d[3] + d[5] * d[5] * lam + 2.0 * d[5] * d[6] * d[0] + d[6] * d[6] * d[1]) * ninv * ninv;
            }
        }
        double a[4];
        double theta = 0.0;
        if (!ritz_max(B, use_p ? 3 : 2, a, &theta)) {     // (every thread sees the same B: the branch is uniform)
            double low = fmin(fmin(B[0][0] - fabs(B[0][1]) - fabs(B[0][2]), B[1][1] - fabs(B[0][1]) - fabs(B[1][2])),
                              B[2][2] - fabs(B[0][2]) - fabs(B[1][2])) - 1.0 - fabs(lam);
            if (!use_p) B[2][2] = low - 1.0;             // decoupled filler rows below the spectrum
            B[3][3] = low - 2.0;
            jacobi4_max(B, a);
            theta = 0.0;
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) theta += a[i] * B[i][j] * a[j];
        }
        if (a[0] < 0.0) { a[0] = -a[0]; a[1] = -a[1]; a[2] = -a[2]; }
        const double c2 = use_p ? a[2] * ninv : 0.0;
        const double c1 = use_p ? a[1] - c2 * d[6] : a[1];
        const double c0 = use_p ? a[0] - c2 * d[5] : a[0];
        for (int c = tid; c < N; c += T) {
            const double xc = x[c], rc = r[c], arc = Ar[c], axc = Ax[c];
            const double pc = use_p ? p[c] : 0.0, apc = use_p ? Ap[c] : 0.0;
            x[c] = c0 * xc + c1 * rc + c2 * pc;
            Ax[c] = c0 * axc + c1 * arc + c2 * apc;
            p[c] = c1 * rc + c2 * pc;
            Ap[c] = c1 * arc + c2 * apc;
        }
        lam = theta;
        have_p = true; fresh = false;
        __syncthreads();
    }
    {   // unit norm (the basis is orthonormal up to rounding)
        double v[1] = {0.0};
        for (int c = tid; c < N; c += T) v[0] += x[c] * x[c];
        block_sum<1>(sh, v, red_buf);
        const double inv = v[0] > 0.0 ? rsqrt(v[0]) : 0.0;
        for (int c = tid; c < N; c += T) x[c] *= inv;
    }
    __syncthreads();
    return it;
}

// Leading eigenvector of A = S (diag(h) W + W diag(h)) S (S = diag(sv) or identity; h lives in pv.res) by power
// iteration -- the compact-matrix equivalent of csc_matrix + eigs(k=1) (rpmodule.py:270-276, :131-136).
// The iterate is kept un-normalised (y_k = A u_k, u_{k+1} = y_k/||y_k||) and the step's single reduction
// carries both ||y_k||^2 and the residual ||y_k - lambda_{k-1} u_k||^2, i.e. lambda^2 ||u_{k+1}-u_k||^2 with a
// one-step-stale lambda.  On return pv.ua holds the unit eigenvector (non-negative).
template <bool USE_SV, bool ROBUST>
__device__ int power_iteration(Shared& sh, const PairView& pv, double tol, int max_it, bool warm,
                               int& red_buf, int* converged, int fast_cap, int pi_switch) {
    const int tid = threadIdx.x;
    const int N = pv.N;
    for (int c = tid; c < N; c += T) pv.ub[c] = 0.0;     // rows without non-zeros stay 0 in both buffers
    if (!warm) {
        double v[1] = {0.0};
        for (int c = tid; c < N; c += T) { double d = pv.geo[G_DEG * pv.gstride + c]; pv.ua[c] = d; v[0] += d * d; }
        block_sum<1>(sh, v, red_buf);
        double inv0 = v[0] > 0.0 ? rsqrt(v[0]) : 0.0;
        for (int c = tid; c < N; c += T) pv.ua[c] *= inv0;
    }
    __syncthreads();
    PowerStep<USE_SV> st;
    st.cur = pv.ua; st.nxt = pv.ub; st.h = pv.res; st.sv = pv.sv;
    st.inv = 1.0; st.lam_prev = 0.0;
    int it = 0;
    *converged = 0;
    double res_prev = CUDART_INF;
    const int cap = ROBUST ? pi_switch : fast_cap;
    const int max_pi = max_it < cap ? max_it : cap;
    bool dead = false;
    for (it = 0; it < max_pi; ++it) {
        st.acc0 = 0.0; st.acc1 = 0.0;
        csr_walk(sh, pv, st);
        double v[2] = {st.acc0, st.acc1};
        block_sum<2>(sh, v, red_buf);                    // barrier: nxt complete, cur no longer read
        double nrm2 = v[0];
        { const double* t = st.cur; st.cur = st.nxt; st.nxt = const_cast<double*>(t); }
        if (!(nrm2 > 0.0)) { st.inv = 0.0; ++it; dead = true; break; }
        double lam = sqrt(nrm2);
        st.inv = 1.0 / lam;
        if (it > 0) {
            double rel2 = v[1] / nrm2;                   // ~ ||u_{k+1} - u_k||^2
            if (rel2 <= tol * tol) { ++it; *converged = 1; break; }
            if (rel2 >= res_prev && rel2 < 1e-26) { ++it; *converged = 1; break; }   // stagnation at rounding level
            res_prev = rel2;
        }
        st.lam_prev = lam;
    }
    { const double* cur = st.cur; const double inv = st.inv;
      for (int p = tid; p < N; p += T) pv.ua[p] = cur[p] * inv; }   // leave the unit vector in pv.ua
    __syncthreads();
    if (ROBUST) {
        if (!*converged && !dead && it >= pi_switch && max_it > pi_switch)
            it += lopcg_continue<USE_SV>(sh, pv, tol, max_it - it, red_buf, converged);
    } else if (dead) {
        *converged = 1;                                  // zero matrix: nothing a second pass could improve
    }
    return it;
}

// xdeg_c = u_c (W u)_c : degrees of x = max(0,u_p u_q) w_pq (rpmodule.py:277-285; u >= 0 for a Perron vector).
__device__ void x_degrees(Shared& sh, const PairView& pv, double mu) {
    for (int p = threadIdx.x; p < pv.N; p += T)
        if (pv.rowstart[p] == pv.rowstart[p + 1]) { pv.aP[p] = 0.0; pv.aN[p] = 0.0; }
    XDegStep st; st.u = pv.ua; st.aP = pv.aP; st.aN = pv.aN; st.mu = mu;
    csr_walk(sh, pv, st);
    __syncthreads();
}

// res <- h = max(0, OFFSET - res)   (rpmodule.py:265-266: a = w*(offset - r), clipped at 0)
__device__ void residual_to_h(const PairView& pv) {
    for (int cr = threadIdx.x; cr < pv.nrows; cr += T) { const int c = pv.rowmap[cr] & RM_ROW; pv.res[c] = fmax(0.0, OFFSET - pv.res[c]); }
    __syncthreads();
}

template <bool NARROW> struct MaskOf { typedef unsigned type; };
template <> struct MaskOf<false> { typedef unsigned long long type; };

// Phase C: conservative float32 pre-test of the distance-consistency condition (rpmodule.py:399-404).
// The source distance depends only on the source keypoint pair (i1,i2) and is shared by the KK*KK correspondence pairs
// built on it.  The n_s(n_s-1)/2 source pairs are enumerated in row-major triangular order, one per lane (every lane of
// every warp step is busy, whatever n_s is): the lane loads both keypoints and their 2*KK target points and runs the KK*KK
// tests, recording its keeps in a KK*KK-bit mask; one warp scan per step then appends them to the per-warp stage buffer,
// which is flushed 32 entries at a time into the candidate list.
template <int KK>
__device__ void pretest_pairs(Shared& sh, const PairView& pv, float tau, float sep2, long long edge_cap) {
    typedef typename MaskOf<(KK * KK <= 32)>::type MaskT;             // one bit per (k1, k2)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ns = pv.ns;
    const float4* sp4 = pv.sp4; const float4* tq4 = pv.tq4;
    unsigned* stg = sh.stage[warp];
    int nst = 0;                                                     // staged entries (warp-uniform)
    const int P = ns * (ns - 1) / 2;
    for (int p0 = warp * 32; p0 < P; p0 += T) {
        {
            const int p = p0 + lane;
            const bool vs0 = p < P;
            const int pc = vs0 ? p : P - 1;
            int i2 = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)pc)) * 0.5f);      // largest i2 with i2 (i2 - 1) / 2 <= pc
            while (i2 * (i2 - 1) / 2 > pc) --i2;
            while ((i2 + 1) * i2 / 2 <= pc) ++i2;
            const int i1 = pc - i2 * (i2 - 1) / 2;                                // 0 <= i1 < i2
            const float4 p1 = sp4[i1], p2 = sp4[i2];
            float4 q2[KK];
#pragma unroll
            for (int k = 0; k < KK; ++k) q2[k] = tq4[i2 * KK + k];
            float ax = p1.x - p2.x, ay = p1.y - p2.y, az = p1.z - p2.z;
            float S = ax * ax + ay * ay + az * az;
            float ds = S * rsqrt_approx(S + 1e-30f);
            const bool vs = vs0 && (S > sep2);                      // min(ds,dt) > sep  =>  ds > sep
            // |ds - dt| < tau and dt > sep as bounds on the SQUARED target distance (no square root per test)
            const float hi2 = (ds + tau) * (ds + tau);
            const float lo2 = fmaxf(ds > tau ? (ds - tau) * (ds - tau) : -1.f, sep2);
            MaskT keepm = 0;
#pragma unroll
            for (int k1 = 0; k1 < KK; ++k1) {
                const float4 q1 = tq4[i1 * KK + k1];
#pragma unroll
                for (int k2 = 0; k2 < KK; ++k2) {
                    float bx = q1.x - q2[k2].x, by = q1.y - q2[k2].y, bz = q1.z - q2[k2].z;
                    float Tq = bx * bx + by * by + bz * bz;
                    bool keep = (Tq > lo2) && (Tq < hi2);
                    keepm |= keep ? ((MaskT)1 << (k1 * KK + k2)) : (MaskT)0;
                }
            }
            if (!vs) keepm = 0;
            const int cntl = __popcll((unsigned long long)keepm);
            if (__any_sync(0xffffffffu, cntl != 0)) {
                int inc = cntl;                                      // inclusive warp scan of the per-lane counts
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += v; }
                const int total = __shfl_sync(0xffffffffu, inc, 31);
                int pos = nst + inc - cntl;
                // stage buffer holds 64: flush first if this cell could overflow it
                if (nst + total > STAGE) {
                    // emit what is staged (nst < 32 entries) to keep the buffer bounded
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&sh.cnt[0], nst);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (lane < nst && (long long)base + lane < edge_cap) pv.edges[base + lane] = stg[lane];
                    __syncwarp();
                    pos -= nst; nst = 0;
                }
                if (total <= STAGE) {
                    while (keepm) {
                        int bit = __ffsll((long long)(unsigned long long)keepm) - 1; keepm &= keepm - 1;
                        int k1 = bit / KK, k2 = bit - k1 * KK;
                        stg[pos++] = ((unsigned)(i1 * KK + k1) << 16) | (unsigned)(i2 * KK + k2);
                    }
                    nst += total;
                    __syncwarp();
                    while (nst >= 32) {                              // flush 32 entries, coalesced
                        int base = 0;
                        if (lane == 0) base = atomicAdd(&sh.cnt[0], 32);
                        base = __shfl_sync(0xffffffffu, base, 0);
                        unsigned v = stg[lane];
                        if ((long long)base + lane < edge_cap) pv.edges[base + lane] = v;
                        __syncwarp();
                        unsigned mv = (lane < nst - 32) ? stg[32 + lane] : 0u;
                        __syncwarp();
                        if (lane < nst - 32) stg[lane] = mv;
                        nst -= 32;
                        __syncwarp();
                    }
                } else {
                    // dense cell (more than 64 keeps among 32*KK*KK tests): write straight to the candidate list
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&sh.cnt[0], total);
                    base = __shfl_sync(0xffffffffu, base, 0);
                    long long o = (long long)base + (pos - nst);
                    while (keepm) {
                        int bit = __ffsll((long long)(unsigned long long)keepm) - 1; keepm &= keepm - 1;
                        int k1 = bit / KK, k2 = bit - k1 * KK;
                        if (o < edge_cap) pv.edges[o] = ((unsigned)(i1 * KK + k1) << 16) | (unsigned)(i2 * KK + k2);
                        ++o;
                    }
                }
            }
        }
    }
    if (nst > 0) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&sh.cnt[0], nst);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (lane < nst && (long long)base + lane < edge_cap) pv.edges[base + lane] = stg[lane];
    }
}

__device__ __forceinline__ void write_identity(double* T_out) {
    if (T_out && threadIdx.x < 16) T_out[threadIdx.x] = ((threadIdx.x % 5) == 0) ? 1.0 : 0.0;
}

// ROBUST = false: the fused kernel with a plain power iteration capped at PI_FAST_CAP steps; a pair whose iteration has
// not converged by then is marked RP_STATUS_RETRY and left.  ROBUST = true: launched right behind it on the same
// stream, redoes exactly those pairs with the accelerated eigen iteration (lopcg_continue) -- a few microseconds when
// there is none, and the common path keeps its register allocation.
// BIG = true: pairs whose vectors do not fit shared memory keep them in the slot's workspace.  A template parameter, not a
// run-time select, so that in the common variant every access to the per-pair vectors is a known shared-memory access.
#define RP_PHASE_CLK(k) do { if (A.has_dbg && A.dbg.phase_clk && tid == 0) A.dbg.phase_clk[(size_t)b * 8 + (k)] = clock64(); } while (0)

template <bool ROBUST, bool BIG>
__global__ void __launch_bounds__(T, RP_MIN_BLOCKS) rp_solve_kernel(const SolveArgs A) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    __shared__ Shared sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int* work_counter = reinterpret_cast<int*>(A.ws) + (ROBUST ? 1 : 0);
    char* slot = A.ws + 256 + (size_t)blockIdx.x * A.slot_bytes;
    // per-pair vectors / staging: dynamic shared memory, or (N too large for it) the same layout in the slot's workspace
    unsigned char* const dyn = BIG ? reinterpret_cast<unsigned char*>(slot + A.o_dyn) : dyn_smem;

    for (;;) {
        __syncthreads();
        if (tid == 0) { sh.cnt[4] = atomicAdd(work_counter, 1); sh.ties = 0; }
        __syncthreads();
        const int b = sh.cnt[4];
        if (b >= A.B) break;
        if (ROBUST && A.status[b] != RP_STATUS_RETRY) continue;

        const rp_params par = A.params[A.param_idx ? A.param_idx[b] : 0];
        const int s0 = A.off_s[b], t0 = A.off_t[b];
        const int ns = A.off_s[b + 1] - s0, nt = A.off_t[b + 1] - t0;
        double* Tout = A.T_out ? A.T_out + (size_t)b * 16 : nullptr;
        int* st = A.stats ? A.stats + (size_t)b * RP_STATS_STRIDE : nullptr;
        if (st && tid < RP_STATS_STRIDE) st[tid] = 0;
        const int D = A.feat_dim;
        RP_PHASE_CLK(0);

        if (!A.solve_only && (ns < 3 || nt < 3)) {                // rpmodule.py:346-348
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_FEW_KEYPOINTS;
            continue;
        }
        int K = par.topk < nt - 1 ? par.topk : nt - 1;           // rpmodule.py:368
        if (A.solve_only) K = 1;                                  // nodes are the correspondences themselves
        if (K > KMAX || K > A.max_topk || K < 1 || ns * K > A.Nmax) {
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_UNSUPPORTED;
            continue;
        }
        PairView pv;
        pv.ns = ns; pv.nt = nt; pv.K = K; pv.N = ns * K; pv.NW = (pv.N + 31) >> 5;
        pv.geo = reinterpret_cast<double*>(slot + A.o_geo); pv.gstride = A.Nmax;
        pv.hgeo = pv.geo; pv.hgs = A.Nmax;
        pv.cj = reinterpret_cast<int*>(slot + A.o_cj);
        pv.edges = reinterpret_cast<unsigned*>(slot + A.o_edges);
        pv.ew = reinterpret_cast<double*>(slot + A.o_ew);
        pv.cols = reinterpret_cast<uint16_t*>(slot + A.o_cols);
        pv.vals = reinterpret_cast<double*>(slot + A.o_vals);
        {
            double* v = reinterpret_cast<double*>(dyn);
            pv.aP = v; pv.aN = v + A.Nmax; pv.res = v + 2 * A.Nmax; pv.ua = v + 3 * A.Nmax;
            pv.ub = v + 4 * A.Nmax; pv.sv = v + 5 * A.Nmax;
            pv.S = v + 6 * A.Nmax;                                   // 2*Nmax doubles
            pv.rowstart = reinterpret_cast<int*>(v + 8 * A.Nmax);    // Nmax+1 ints
            pv.rowmap = reinterpret_cast<unsigned*>(pv.rowstart + (A.Nmax + 1));   // Nmax words
            pv.E = 1; pv.nnz = 0; pv.nrows = 0;
            pv.cols_s = pv.cols; pv.vals_s = pv.vals; pv.ism = 0;
            pv.sp4 = reinterpret_cast<float4*>(dyn);
            pv.tq4 = pv.sp4 + ns;
            pv.mask = A.mask_in_smem ? reinterpret_cast<unsigned*>(dyn + A.sm_mask_off)
                                     : reinterpret_cast<unsigned*>(slot + A.o_mask);
        }
        const int N = pv.N, NW = pv.NW;
        if (st && tid == 0) { st[0] = N; st[7] = K; }

        // ------------------------------------------------------------------ A. descriptor front end
        if (!A.solve_only) {
            float* tfeat = reinterpret_cast<float*>(dyn);            // aliases the vectors (not yet live)
            const int ts = A.tfeat_stride;
            const bool vec4 = ((D & 7) == 0) && ((ts & 3) == 0);
            const bool seq_sum = A.feat_sum_order && A.feat_sum_order[b] != 0;
            const float* ft = A.feat_t + (size_t)t0 * D;
            for (int e = tid; e < nt * D; e += T) {
                int j = e / D, c = e - j * D;
                tfeat[j * ts + c] = __fdiv_rn(ft[e], FEAT_SCALING);    // rpmodule.py:343
            }
            __syncthreads();
            if (D == 32 && vec4 && (reinterpret_cast<uintptr_t>(A.feat_s) & 15) == 0) {
                // One source keypoint per THREAD: its 32 channels live in registers, the target rows are read from shared
                // memory as warp-wide broadcasts, and distance, soft-match norm and the top-k list are all thread-local --
                // no shuffles, no per-(i,j) division (candidates are ranked by dij * (1/den), a monotone image of the
                // reference's exp(-dij/den); the exact quotient is only formed for entries that reach the norm or the list).
                // The 512-thread build gives a source keypoint QA = 4 adjacent lanes, each scanning a quarter of the targets; the
                // four sorted lists are merged with a shuffle butterfly under the same (key, index) order the sequential scan
                // produces, and the norm is the four partial sums added in lane order.
                const double rden_obs = 1.0 / par.feat_den_obs, rden_any = 1.0 / par.feat_den;
                constexpr int QA = T >= 512 ? 4 : 1;
                constexpr int RPP = T / QA;
                for (int i0 = 0; i0 < ns; i0 += RPP) {
                    const int i = i0 + tid / QA, q = tid % QA;
                    const bool act = i < ns;
                    const int ic = act ? i : ns - 1;
                    unsigned long long sreg[16];
                    {
                        const float4* fs4 = reinterpret_cast<const float4*>(A.feat_s + (size_t)(s0 + ic) * 32);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 v = fs4[q];
                            const float a0 = __fdiv_rn(v.x, FEAT_SCALING), a1 = __fdiv_rn(v.y, FEAT_SCALING);      // :342
                            const float a2 = __fdiv_rn(v.z, FEAT_SCALING), a3 = __fdiv_rn(v.w, FEAT_SCALING);
                            sreg[2 * q] = (unsigned long long)__float_as_uint(a0) | ((unsigned long long)__float_as_uint(a1) << 32);
                            sreg[2 * q + 1] = (unsigned long long)__float_as_uint(a2) | ((unsigned long long)__float_as_uint(a3) << 32);
                        }
                    }
                    const double wsi = A.w_s[s0 + ic];
                    double lk[KMAX]; int li[KMAX];
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) { lk[k] = CUDART_INF; li[k] = 0x7fffffff; }
                    double ss = 0.0;
                    const int jq0 = QA == 1 ? 0 : (nt * q) / QA, jq1 = QA == 1 ? nt : (nt * (q + 1)) / QA;
                    for (int j = jq0; j < jq1; ++j) {
                        float sq[32];
                        sqdiff32(sreg, tfeat + j * ts, sq);
                        const float dij = seq_sum ? sum32_seq(sq) : sum32_numpy(sq);                       // :355
                        if (A.has_dbg && A.dbg.dij && act) A.dbg.dij[A.dbg.dij_off[b] + (int64_t)i * nt + j] = dij;
                        const bool obs = __dmul_rn(wsi, __ldg(A.w_t + t0 + j)) == 1.0;                     // :354
                        const double qk = (double)dij * (obs ? rden_obs : rden_any);                       // ranks like -key
                        // norm terms: below -372.6 exp(key)^2 < 2^-1075 rounds to +0; a term more than 22 below the best key seen
                        // so far is < e^-44 of a term already in the sum (n_t of them stay under 1e-17 relative): both are no-ops
                        if (qk < 372.6000001 && qk < lk[0] + 22.0) {
                            const double key = (double)(-dij) / (obs ? par.feat_den_obs : par.feat_den);   // :356-358
                            if (key > -372.6) { const double e = exp(key); ss += e * e; }
                        }
                        if (qk < lk[KMAX - 1]) {               // strict: of equal keys the lower index stays ahead
                            double ck = qk; int ci = j;
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) {
                                if (ck < lk[k]) { const double tk = lk[k]; const int ti = li[k]; lk[k] = ck; li[k] = ci; ck = tk; ci = ti; }
                            }
                        }
                    }
                    if (QA > 1) {
#pragma unroll
                        for (int x = 1; x < QA; x <<= 1) {
                            double ok[KMAX]; int oi[KMAX];
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) { ok[k] = shfl_xor_d(lk[k], x); oi[k] = __shfl_xor_sync(0xffffffffu, li[k], x); }
#pragma unroll
                            for (int k2 = 0; k2 < KMAX; ++k2) {
                                double ck = ok[k2]; int ci = oi[k2];
                                if (ck < lk[KMAX - 1] || (ck == lk[KMAX - 1] && ci < li[KMAX - 1])) {
#pragma unroll
                                    for (int k = 0; k < KMAX; ++k) {
                                        if (ck < lk[k] || (ck == lk[k] && ci < li[k])) { const double tk = lk[k]; const int ti = li[k]; lk[k] = ck; li[k] = ci; ck = tk; ci = ti; }
                                    }
                                }
                            }
                        }
                        double tot = 0.0;
#pragma unroll
                        for (int qq = 0; qq < QA; ++qq) tot += __shfl_sync(0xffffffffu, ss, (lane & ~(QA - 1)) + qq);
                        ss = tot;
                    }
                    if (act) {
                        const double nm = sqrt(ss);                                                        // :359
                        // The index SET is decided by the keys unless the K-th and (K+1)-th candidate tie exactly, or selected
                        // entries have underflowed to wij = 0 in a row that is not all zero (exp(key) = 0 below -745.1): there
                        // numpy's introselect order decides, which is not reproduced (all-zero rows are: zero_row_topk).
                        if (nm != 0.0) {
                            bool amb = false;
#pragma unroll
                            for (int k = 1; k < KMAX; ++k) if (k == K && lk[k] == lk[k - 1] && li[k] != 0x7fffffff) amb = true;
#pragma unroll
                            for (int k = 0; k < KMAX; ++k) if (k == K - 1 && lk[k] > 745.13 && li[k] != 0x7fffffff) amb = true;
                            if (amb && q == 0) atomicAdd(&sh.ties, 1);
                        }
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) {
                            if ((k % QA) != q) continue;                          // the row's output entries are dealt over its lanes
                            if (k < K) {
                                int idx = li[k];
                                double f = 0.0;
                                if (idx == 0x7fffffff) idx = k < nt ? k : 0;      // NaN descriptors: arbitrary but valid
                                else if (nm != 0.0) {
                                    float sq[32];
                                    sqdiff32(sreg, tfeat + idx * ts, sq);
                                    const float dk = seq_sum ? sum32_seq(sq) : sum32_numpy(sq);
                                    const bool obs = __dmul_rn(wsi, __ldg(A.w_t + t0 + idx)) == 1.0;
                                    f = exp((double)(-dk) / (obs ? par.feat_den_obs : par.feat_den)) / nm;  // :360-363
                                }
                                if (nm == 0.0 && A.zero_row_topk) {                // all-zero row: numpy's tie order (host supplied)
                                    const int z = A.zero_row_topk[(size_t)b * A.max_topk + k];
                                    if (z >= 0 && z < nt) idx = z;
                                }
                                const int c = i * K + k;
                                pv.cj[c] = idx;
                                pv.geo[G_F * pv.gstride + c] = f;
                                if (A.has_dbg && A.dbg.topk_idx) A.dbg.topk_idx[(size_t)(s0 + i) * A.max_topk + k] = idx;
                                if (A.has_dbg && A.dbg.topk_f) A.dbg.topk_f[(size_t)(s0 + i) * A.max_topk + k] = f;
                            } else if (k < A.max_topk) {
                                if (A.has_dbg && A.dbg.topk_idx) A.dbg.topk_idx[(size_t)(s0 + i) * A.max_topk + k] = -1;
                                if (A.has_dbg && A.dbg.topk_f) A.dbg.topk_f[(size_t)(s0 + i) * A.max_topk + k] = 0.0;
                            }
                        }
                    }
                }
            } else
            for (int i = warp; i < ns; i += NWARP) {
                const float* fs = A.feat_s + (size_t)(s0 + i) * D;
                for (int c = lane; c < D; c += 32) sh.sfeat[warp][c] = __fdiv_rn(fs[c], FEAT_SCALING);   // :342
                __syncwarp();
                const double wsi = A.w_s[s0 + i];
                double lk[KMAX]; int li[KMAX];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) { lk[k] = -CUDART_INF; li[k] = 0x7fffffff; }
                double ss = 0.0;
                for (int j = lane; j < nt; j += 32) {
                    float dij = seq_sum ? numpy_sqdist_f32_seq(sh.sfeat[warp], tfeat + j * ts, D)
                                : vec4  ? numpy_sqdist_f32_v4(sh.sfeat[warp], tfeat + j * ts, D)
                                        : numpy_sqdist_f32(sh.sfeat[warp], tfeat + j * ts, D);     // :355
                    if (A.has_dbg && A.dbg.dij) A.dbg.dij[A.dbg.dij_off[b] + (int64_t)i * nt + j] = dij;
                    double both = __dmul_rn(wsi, A.w_t[t0 + j]);                                   // :354
                    double den = (both == 1.0) ? par.feat_den_obs : par.feat_den;                 // :356-357
                    double key = (double)(-dij) / den;                                            // :358 (argument of exp)
                    if (key > -372.6) {          // below this exp(key)^2 < 2^-1075 rounds to +0: adding it is a no-op
                        double e = exp(key);
                        ss += e * e;
                    }
                    if (key > lk[KMAX - 1]) {
                        double ck = key; int ci = j;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) {
                            if (ck > lk[k]) { double tk = lk[k]; int ti = li[k]; lk[k] = ck; li[k] = ci; ck = tk; ci = ti; }
                        }
                    }
                }
                ss = warp_sum(ss);
                double nm = sqrt(ss);                                                              // :359
                double mykey = 0.0; int myidx = -1;
                for (int r = 0; r < K; ++r) {
                    double bk = lk[0]; int bi = li[0];
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        double ok = shfl_xor_d(bk, o); int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                        if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
                    }
                    if (li[0] == bi && bi != 0x7fffffff) {      // this lane owned the winner: pop it
#pragma unroll
                        for (int k = 0; k < KMAX - 1; ++k) { lk[k] = lk[k + 1]; li[k] = li[k + 1]; }
                        lk[KMAX - 1] = -CUDART_INF; li[KMAX - 1] = 0x7fffffff;
                    }
                    if (lane == r) { mykey = bk; myidx = bi; }
                }
                if (lane < K) {
                    if (myidx == 0x7fffffff || myidx < 0) myidx = lane < nt ? lane : 0;   // NaN descriptors: arbitrary but valid
                    double f = (nm == 0.0) ? 0.0 : exp(mykey) / nm;                               // :360-363
                    if (nm == 0.0 && A.zero_row_topk) {         // all-zero row: numpy's tie order (host supplied)
                        int z = A.zero_row_topk[(size_t)b * A.max_topk + lane];
                        if (z >= 0 && z < nt) myidx = z;
                    }
                    int c = i * K + lane;
                    pv.cj[c] = myidx;
                    pv.geo[G_F * pv.gstride + c] = f;
                    if (A.has_dbg && A.dbg.topk_idx) A.dbg.topk_idx[(size_t)(s0 + i) * A.max_topk + lane] = myidx;
                    if (A.has_dbg && A.dbg.topk_f) A.dbg.topk_f[(size_t)(s0 + i) * A.max_topk + lane] = f;
                } else if (lane < A.max_topk) {
                    if (A.has_dbg && A.dbg.topk_idx) A.dbg.topk_idx[(size_t)(s0 + i) * A.max_topk + lane] = -1;
                    if (A.has_dbg && A.dbg.topk_f) A.dbg.topk_f[(size_t)(s0 + i) * A.max_topk + lane] = 0.0;
                }
                __syncwarp();
            }
            __syncthreads();
        }
        RP_PHASE_CLK(1);
        if (A.stop_after == RP_STAGE_TOPK) {
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_OK;
            continue;
        }
        if (!A.solve_only && N < 3) {                                // rpmodule.py:377-379
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_FEW_CORRES;
            continue;
        }

        // ------------------------------------------------------------------ B. per-correspondence geometry
        int red_buf = 0;
        {
            const int gs = pv.gstride;
            double amax = 0.0;
            for (int c = tid; c < N; c += T) {
                int i = c / K, j = A.solve_only ? c : pv.cj[c];      // solve-only: node c carries both sides (t0 == s0)
                const double* p = A.pc_s + (size_t)(s0 + i) * 3; const double* q = A.pc_t + (size_t)(t0 + j) * 3;
                const double* n = A.nrm_s + (size_t)(s0 + i) * 3; const double* m = A.nrm_t + (size_t)(t0 + j) * 3;
                double p0 = p[0], p1 = p[1], p2 = p[2], q0 = q[0], q1 = q[1], q2 = q[2];
                pv.geo[G_PX * gs + c] = p0; pv.geo[G_PY * gs + c] = p1; pv.geo[G_PZ * gs + c] = p2;
                pv.geo[G_QX * gs + c] = q0; pv.geo[G_QY * gs + c] = q1; pv.geo[G_QZ * gs + c] = q2;
                pv.geo[G_NX * gs + c] = n[0]; pv.geo[G_NY * gs + c] = n[1]; pv.geo[G_NZ * gs + c] = n[2];
                pv.geo[G_MX * gs + c] = m[0]; pv.geo[G_MY * gs + c] = m[1]; pv.geo[G_MZ * gs + c] = m[2];
                if (!A.solve_only) { pv.geo[G_WS * gs + c] = A.w_s[s0 + i]; pv.geo[G_WT * gs + c] = A.w_t[t0 + j]; }
                pv.tq4[c] = make_float4((float)q0, (float)q1, (float)q2, 0.f);
                if (c - i * K == 0) pv.sp4[i] = make_float4((float)p0, (float)p1, (float)p2, 0.f);
                amax = fmax(amax, fmax(fmax(fabs(p0), fabs(p1)), fmax(fabs(p2), fmax(fabs(q0), fmax(fabs(q1), fabs(q2))))));
            }
            for (int e = tid; e < N * NW; e += T) pv.mask[e] = 0u;
            if (tid < 8) sh.cnt[tid] = (tid == 4) ? b : 0;
            if (tid == 0) { sh.scal[0] = 0.0; sh.scal[2] = 0.0; sh.wmax_bits = 0ull; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amax = fmax(amax, shfl_xor_d(amax, o));
            if (lane == 0) sh.red[red_buf][warp][0] = amax;
            __syncthreads();
            if (tid == 0) {
                double m = 0.0;
                for (int w = 0; w < NWARP; ++w) m = fmax(m, sh.red[red_buf][w][0]);
                // float32 pre-test margin: coordinates rounded to f32 (rel 2^-24), differences, squares, sqrt --
                // absolute distance error < ~8 * 2^-24 * max|coord| ~ 5e-7 * max|coord|; 1e-3 floor.  See DESIGN.md.
                sh.scal[1] = 1.0e-3 + 4.0e-6 * m;
            }
            red_buf ^= 1;
            __syncthreads();
        }

        // ------------------------------------------------------------------ C'/D'. solve-only: import the caller's pair list
        if (A.solve_only) {
            const int e0 = A.edge_off ? A.edge_off[b] : 0, ne = A.edge_off ? A.edge_off[b + 1] - e0 : 0;
            if ((long long)ne > A.edge_cap) {
                write_identity(Tout);
                if (tid == 0) A.status[b] = RP_STATUS_EDGE_OVERFLOW;
                continue;
            }
            for (int e = tid; e < ne; e += T) {
                int r = A.edge_rc[2 * (size_t)(e0 + e)], c = A.edge_rc[2 * (size_t)(e0 + e) + 1];
                if (r > c) { int t2 = r; r = c; c = t2; }
                const bool ok = r >= 0 && c < N && r != c;
                pv.edges[e] = ((unsigned)(ok ? r : 0) << 16) | (unsigned)(ok ? c : 0);
                pv.ew[e] = ok ? A.edge_w[e0 + e] : -1.0;
                if (ok) {
                    atomicOr(&pv.mask[(size_t)r * NW + (c >> 5)], 1u << (c & 31));
                    atomicOr(&pv.mask[(size_t)c * NW + (r >> 5)], 1u << (r & 31));
                }
            }
            if (tid == 0) { sh.cnt[0] = ne; sh.cnt[1] = ne; sh.cnt[2] = ne; sh.cnt[5] = ne; }
            __syncthreads();
        }
        // ------------------------------------------------------------------ C. float32 pre-test -> candidates  (:389-404)
        if (!A.solve_only) {
            const float margin = (float)sh.scal[1];
            const float tau = (float)sqrt(par.dist_thre_sq) + margin;       // |ds - dt| < distThre (+margin)
            float sep = (float)par.sep_thre - margin;                       // min(ds,dt) > 1.5*distSepThre^2 (-margin)
            const float sep2 = sep > 0.f ? sep * sep : -1.f;                // compared against squared distances
            switch (K) {
                case 1: pretest_pairs<1>(sh, pv, tau, sep2, A.edge_cap); break;
                case 2: pretest_pairs<2>(sh, pv, tau, sep2, A.edge_cap); break;
                case 3: pretest_pairs<3>(sh, pv, tau, sep2, A.edge_cap); break;
                case 4: pretest_pairs<4>(sh, pv, tau, sep2, A.edge_cap); break;
                case 5: pretest_pairs<5>(sh, pv, tau, sep2, A.edge_cap); break;
                case 6: pretest_pairs<6>(sh, pv, tau, sep2, A.edge_cap); break;
                case 7: pretest_pairs<7>(sh, pv, tau, sep2, A.edge_cap); break;
                default: pretest_pairs<8>(sh, pv, tau, sep2, A.edge_cap); break;
            }
            __syncthreads();
        }
        const int MC = sh.cnt[0];
        RP_PHASE_CLK(2);
        if ((long long)MC > A.edge_cap) {
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_EDGE_OVERFLOW;
            continue;
        }

        // ------------------------------------------------------------------ D. exact tests + pair weight (:399-467)
        // Stage 1, every candidate: the exact float64 distance test in NumPy's operation order (counts M1), then a cheap
        // rejection of the three angle conditions -- |acos x - acos y| < th  <=>  x y + sqrt((1-x^2)(1-y^2)) > cos th on
        // [0, pi], one square root instead of two acos -- with a 1e-9 margin that can only let a pair THROUGH.  Candidates
        // that are not certainly rejected (a fifth of them) are compacted into a survivor list.  Stage 2, survivors only,
        // all lanes busy: the reference's own arithmetic (six acos, the filters exactly as :430-436, the pair weight).
        if (!A.solve_only) {
            const int gs = pv.gstride; const double* geo = pv.geo;
            unsigned* surv = reinterpret_cast<unsigned*>(slot + A.o_cols);   // the CSR arrays are not built yet
            int m1 = 0, m2 = 0, nz = 0;
            double wloc = 0.0;
            const double th = sqrt(par.angle_thre_sq);
            const bool cheap = th > 0.0 && th < 3.14159;                // otherwise everything goes to stage 2
            // the rejection runs in float32 on float32 copies of the unit normals: rounding of the dot products and of
            // sqrt((1-x^2)(1-y^2)) near |x| = 1 stays below 6e-4, hence the 2e-3 margin (0.16 degrees at th = 45 degrees)
            const float cth = cheap ? (float)cos(th) - 2e-3f : -4.0f;
            // stage-1 operands in shared memory (the per-pair vectors are not live yet; 72 N bytes always fit their region)
            double* sP = reinterpret_cast<double*>(dyn);                // [6][N] source / target position
            float* sN = reinterpret_cast<float*>(sP + 6 * (size_t)N);   // [6][N] source / target normal
            for (int c = tid; c < N; c += T) {
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    sP[(size_t)a * N + c] = geo[(size_t)(G_PX + a) * gs + c];
                    sN[(size_t)a * N + c] = (float)geo[(size_t)(G_NX + a) * gs + c];
                }
            }
            __syncthreads();
            for (int e0 = 0; e0 < MC; e0 += T) {
                const int e = e0 + tid;
                bool maybe = false;
                if (e < MC) {
                    unsigned rc = pv.edges[e];
                    int r = rc >> 16, c = rc & 0xffffu;
                    double ax = sP[r] - sP[c], ay = sP[N + r] - sP[N + c], az = sP[2 * N + r] - sP[2 * N + c];
                    double bx = sP[3 * N + r] - sP[3 * N + c], by = sP[4 * N + r] - sP[4 * N + c], bz = sP[5 * N + r] - sP[5 * N + c];
                    double dis_s = sqrt(dot3_np(ax, ay, az, ax, ay, az));                // :399
                    double dis_t = sqrt(dot3_np(bx, by, bz, bx, by, bz));                // :400
                    double df = dis_s - dis_t;
                    double dd = __dmul_rn(df, df);                                       // :401
                    if ((dd < par.dist_thre_sq) && (fmin(dis_s, dis_t) > par.sep_thre)) {   // :404
                        ++m1;
                        maybe = true;
                        const float n1x = sN[r], n1y = sN[N + r], n1z = sN[2 * N + r], n2x = sN[c], n2y = sN[N + c], n2z = sN[2 * N + c];
                        const float m1x = sN[3 * N + r], m1y = sN[4 * N + r], m1z = sN[5 * N + r], m2x = sN[3 * N + c], m2y = sN[4 * N + c], m2z = sN[5 * N + c];
                        auto cosdiff = [](float x, float y) {          // cos(acos x - acos y)
                            x = fminf(fmaxf(x, -1.f), 1.f); y = fminf(fmaxf(y, -1.f), 1.f);
                            return x * y + sqrtf((1.f - x * x) * (1.f - y * y));
                        };
                        if (cosdiff(n1x * n2x + n1y * n2y + n1z * n2z, m1x * m2x + m1y * m2y + m1z * m2z) < cth) maybe = false;
                        else {
                            const float is = 1.0f / (float)dis_s, it = 1.0f / (float)dis_t;
                            const float e1x = (float)ax * is, e1y = (float)ay * is, e1z = (float)az * is;
                            const float e2x = (float)bx * it, e2y = (float)by * it, e2z = (float)bz * it;
                            if (cosdiff(n1x * e1x + n1y * e1y + n1z * e1z, m1x * e2x + m1y * e2y + m1z * e2z) < cth) maybe = false;
                            else if (cosdiff(n2x * e1x + n2y * e1y + n2z * e1z, m2x * e2x + m2y * e2y + m2z * e2z) < cth) maybe = false;
                        }
                    }
                    if (!maybe) pv.ew[e] = -1.0;
                }
                const unsigned bal = __ballot_sync(0xffffffffu, maybe);
                if (bal) {
                    int base = 0;
                    if (lane == 0) base = atomicAdd(&sh.cnt[7], __popc(bal));
                    base = __shfl_sync(0xffffffffu, base, 0);
                    if (maybe) surv[base + __popc(bal & ((1u << lane) - 1u))] = (unsigned)e;
                }
            }
            __syncthreads();
            const int NSV = sh.cnt[7];
            for (int sidx = tid; sidx < NSV; sidx += T) {
                const int e = (int)surv[sidx];
                unsigned rc = pv.edges[e];
                int r = rc >> 16, c = rc & 0xffffu;
                double ax = geo[G_PX * gs + r] - geo[G_PX * gs + c], ay = geo[G_PY * gs + r] - geo[G_PY * gs + c], az = geo[G_PZ * gs + r] - geo[G_PZ * gs + c];
                double bx = geo[G_QX * gs + r] - geo[G_QX * gs + c], by = geo[G_QY * gs + r] - geo[G_QY * gs + c], bz = geo[G_QZ * gs + r] - geo[G_QZ * gs + c];
                double dis_s = sqrt(dot3_np(ax, ay, az, ax, ay, az));                // :399
                double dis_t = sqrt(dot3_np(bx, by, bz, bx, by, bz));                // :400
                double df = dis_s - dis_t;
                double dd = __dmul_rn(df, df);                                       // :401
                double w = -1.0;
                {
                    const double is = 1.0 / dis_s, it = 1.0 / dis_t;                    // :424-427 (a * (1/|a|), <= 1 ulp from a/|a|)
                    double e1x = ax * is, e1y = ay * is, e1z = az * is;
                    double e2x = bx * it, e2y = by * it, e2z = bz * it;
                    double n1x = geo[G_NX * gs + r], n1y = geo[G_NY * gs + r], n1z = geo[G_NZ * gs + r];
                    double n2x = geo[G_NX * gs + c], n2y = geo[G_NY * gs + c], n2z = geo[G_NZ * gs + c];
                    double m1x = geo[G_MX * gs + r], m1y = geo[G_MY * gs + r], m1z = geo[G_MZ * gs + r];
                    double m2x = geo[G_MX * gs + c], m2y = geo[G_MY * gs + c], m2z = geo[G_MZ * gs + c];
                    double a1 = acos(clip1(dot3_np(n1x, n1y, n1z, n2x, n2y, n2z))) - acos(clip1(dot3_np(m1x, m1y, m1z, m2x, m2y, m2z)));
                    double b1 = acos(clip1(dot3_np(n1x, n1y, n1z, e1x, e1y, e1z))) - acos(clip1(dot3_np(m1x, m1y, m1z, e2x, e2y, e2z)));
                    double g1 = acos(clip1(dot3_np(n2x, n2y, n2z, e1x, e1y, e1z))) - acos(clip1(dot3_np(m2x, m2y, m2z, e2x, e2y, e2z)));
                    double alpha = __dmul_rn(a1, a1), beta = __dmul_rn(b1, b1), gamma = __dmul_rn(g1, g1);   // :430-432
                    if ((alpha < par.angle_thre_sq) && (beta < par.angle_thre_sq) && (gamma < par.angle_thre_sq)) {   // :434-436
                        double ex = ((-dd / par.den_dist - alpha / par.den_a1) - beta / par.den_a2) - gamma / par.den_a2;   // :457-460
                        w = (geo[G_F * gs + r] * geo[G_F * gs + c]) * exp(ex);
                        double seen = __dmul_rn(__dmul_rn(__dmul_rn(geo[G_WS * gs + r], geo[G_WS * gs + c]), geo[G_WT * gs + r]), geo[G_WT * gs + c]);   // :462-466
                        if (seen != 1.0) w *= UNOBS_DAMP;                                                                      // :467
                        ++m2; if (w != 0.0) ++nz;
                        wloc = fmax(wloc, w);
                    }
                }
                pv.ew[e] = w;
            }
            m1 = __reduce_add_sync(0xffffffffu, m1); m2 = __reduce_add_sync(0xffffffffu, m2); nz = __reduce_add_sync(0xffffffffu, nz);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wloc = fmax(wloc, shfl_xor_d(wloc, o));
            if (lane == 0) {
                atomicAdd(&sh.cnt[1], m1); atomicAdd(&sh.cnt[2], m2); atomicAdd(&sh.cnt[5], nz);
                atomicMax(&sh.wmax_bits, (unsigned long long)__double_as_longlong(wloc));
            }
            __syncthreads();
            // second, light pass: pairs that matter numerically enter the symmetric bit mask (-> CSR)
            const double thr = PRUNE_REL * __longlong_as_double((long long)sh.wmax_bits);
            if (tid == 0) sh.scal[2] = thr;
            for (int e = tid; e < MC; e += T) {
                const double w = pv.ew[e];
                if (w >= thr && w >= 0.0) {
                    unsigned rc = pv.edges[e];
                    int r = rc >> 16, c = rc & 0xffffu;
                    atomicOr(&pv.mask[(size_t)r * NW + (c >> 5)], 1u << (c & 31));
                    atomicOr(&pv.mask[(size_t)c * NW + (r >> 5)], 1u << (r & 31));
                }
            }
            __syncthreads();
        }
        const int M1 = sh.cnt[1], M2 = sh.cnt[2], NZ = sh.cnt[5];
        RP_PHASE_CLK(3);
        if (st && tid == 0) { st[1] = M1; st[2] = M2; st[3] = NZ; }
        if (A.has_dbg && A.dbg.edge_rc) {
            // compacted dump of the surviving pairs (test hook; order follows the candidate list)
            if (tid == 0) {
                long long o = 0;
                for (int e = 0; e < MC; ++e) {
                    if (pv.ew[e] >= 0.0 && o < A.dbg.edge_cap) {
                        unsigned rc = pv.edges[e];
                        A.dbg.edge_rc[((size_t)b * A.dbg.edge_cap + o) * 2 + 0] = rc >> 16;
                        A.dbg.edge_rc[((size_t)b * A.dbg.edge_cap + o) * 2 + 1] = rc & 0xffffu;
                        if (A.dbg.edge_w) A.dbg.edge_w[(size_t)b * A.dbg.edge_cap + o] = pv.ew[e];
                        ++o;
                    }
                }
            }
            __syncthreads();
        }
        if (!A.solve_only && M1 < 3) {                               // rpmodule.py:406-408
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_FEW_DIST;
            continue;
        }
        if (!A.solve_only && M2 < 3) {                               // rpmodule.py:440-443
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_FEW_ANGLE;
            continue;
        }
        if (!A.solve_only && NZ < 1) {                               // rpmodule.py:469-472
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_ZERO_WEIGHT;
            continue;
        }
        if (A.stop_after == RP_STAGE_AFFINITY) {
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_OK;
            continue;
        }

        // ------------------------------------------------------------------ E. CSR of W (both directions)
        {
            int* cidx = reinterpret_cast<int*>(pv.S);                 // compact index of every row (temporary)
            uint16_t* wpre = reinterpret_cast<uint16_t*>(slot + A.o_rowstart);   // popcount of the mask words before word w of row r
            if (tid == 0) { sh.cnt[3] = 0; sh.cnt[6] = 0; }
            __syncthreads();
            for (int base = 0; base < N; base += T) {                 // exclusive scans: row populations, non-empty rows
                int r = base + tid;
                int cnt = 0;
                if (r < N) {
                    uint16_t* wp = wpre + (size_t)r * NW;
                    for (int w = 0; w < NW; ++w) { wp[w] = (uint16_t)cnt; cnt += __popc(pv.mask[(size_t)r * NW + w]); }
                }
                int ne = cnt > 0 ? 1 : 0;
                int inc = cnt, inc2 = ne;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    int v = __shfl_up_sync(0xffffffffu, inc, o), v2 = __shfl_up_sync(0xffffffffu, inc2, o);
                    if (lane >= o) { inc += v; inc2 += v2; }
                }
                if (lane == 31) { sh.warp_tot[warp] = inc; sh.warp_tot2[warp] = inc2; }
                __syncthreads();
                int woff = 0, woff2 = 0;
#pragma unroll
                for (int w = 0; w < NWARP; ++w) if (w < warp) { woff += sh.warp_tot[w]; woff2 += sh.warp_tot2[w]; }
                int carry = sh.cnt[3], carry2 = sh.cnt[6];
                if (r < N) { pv.rowstart[r] = carry + woff + inc - cnt; cidx[r] = carry2 + woff2 + inc2 - ne; }
                __syncthreads();
                if (tid == T - 1) { sh.cnt[3] = carry + woff + inc; sh.cnt[6] = carry2 + woff2 + inc2; }
                __syncthreads();
            }
            if (tid == 0) pv.rowstart[N] = sh.cnt[3];
            __syncthreads();
            pv.nnz = pv.rowstart[N];
            pv.nrows = sh.cnt[6];
            pv.E = (pv.nnz + T - 1) / T; if (pv.E < 1) pv.E = 1;
            const int E = pv.E;
            if (A.csr_smem_cap >= T * CSR_PB) {                     // shared memory an idle SM has left (small batches): the leading
                const int rows_s = A.csr_smem_cap / T;               // entries of every thread's chunk live there, whole batches only
                pv.ism = rows_s >= E ? E + CSR_PB : rows_s / CSR_PB * CSR_PB;
                pv.cols_s = reinterpret_cast<uint16_t*>(dyn_smem + A.sm_csr_off);
                pv.vals_s = reinterpret_cast<double*>(dyn_smem + A.sm_csr_off + (size_t)A.csr_smem_cap * sizeof(uint16_t));
            }
            {   // compact row of this thread's first logical non-zero (upper_bound - 1 over rowstart)
                int k0 = tid * E, lo_r = 0, hi_r = N;                 // invariant: rowstart[lo_r] <= k0 < rowstart[hi_r]
                if (k0 < pv.nnz) {
                    while (hi_r - lo_r > 1) { int mid = (lo_r + hi_r) >> 1; if (pv.rowstart[mid] <= k0) lo_r = mid; else hi_r = mid; }
                    sh.crow0[tid] = cidx[lo_r];
                } else sh.crow0[tid] = 0;
            }
            for (int r = tid; r < N; r += T) {
                int rs = pv.rowstart[r], re = pv.rowstart[r + 1];
                if (re > rs) pv.rowmap[cidx[r]] = (unsigned)r | ((unsigned)(rs / E) << RM_C0) | ((unsigned)((re - 1) / E) << RM_C1);
                else pv.geo[G_DEG * pv.gstride + r] = 0.0;
            }
            const double csr_thr = sh.scal[2];
            for (int e = tid; e < MC; e += T) {
                double w = pv.ew[e];
                if (w < 0.0 || w < csr_thr) continue;
                unsigned rc = pv.edges[e];
                int r = rc >> 16, c = rc & 0xffffu;
                const int rs_r = pv.rowstart[r], rs_c = pv.rowstart[c];
                const int pr = rs_r + wpre[(size_t)r * NW + (c >> 5)] + __popc(pv.mask[(size_t)r * NW + (c >> 5)] & ((1u << (c & 31)) - 1u));
                const int pc = rs_c + wpre[(size_t)c * NW + (r >> 5)] + __popc(pv.mask[(size_t)c * NW + (r >> 5)] & ((1u << (r & 31)) - 1u));
                unsigned fr_flags = (pr == rs_r ? 0x8000u : 0u) | (pr == pv.rowstart[r + 1] - 1 ? 0x4000u : 0u);
                unsigned fc_flags = (pc == rs_c ? 0x8000u : 0u) | (pc == pv.rowstart[c + 1] - 1 ? 0x4000u : 0u);
                const int ir = pr % E, ic = pc % E;
                size_t fr = (size_t)ir * T + pr / E, fc = (size_t)ic * T + pc / E;   // lane-interleaved layout
                (ir < pv.ism ? pv.cols_s : pv.cols)[fr] = (uint16_t)((unsigned)c | fr_flags); (ir < pv.ism ? pv.vals_s : pv.vals)[fr] = w;
                (ic < pv.ism ? pv.cols_s : pv.cols)[fc] = (uint16_t)((unsigned)r | fc_flags); (ic < pv.ism ? pv.vals_s : pv.vals)[fc] = w;
            }
            __syncthreads();
            { DegStep dg; dg.deg = pv.geo + (size_t)G_DEG * pv.gstride; csr_walk(sh, pv, dg); }   // row degrees of W
            __syncthreads();
        }

        if (A.solve_only && A.node_wp) {        // explicit per-node weights (no pair list): every node is active
            for (int c = tid; c < N; c += T) pv.rowmap[c] = (unsigned)c;
            pv.nrows = N;
            __syncthreads();
        }

        RP_PHASE_CLK(4);
        // ------------------------------------------------------------------ F. fitters
        if (A.sm_geo_off) {          // a whole SM per pair: the 60 fit / residual passes read their geometry from shared memory
            double* hg = reinterpret_cast<double*>(dyn_smem + A.sm_geo_off);
            for (int e = tid; e < 12 * N; e += T) { const int a = e / N, c = e - a * N; hg[(size_t)a * N + c] = pv.geo[(size_t)a * pv.gstride + c]; }
            pv.hgeo = hg; pv.hgs = N;
            __syncthreads();
        }
        const double mu = par.mu;
        int tot_it = 0, max_it_seen = 0, not_conv = 0;
        bool retry = false;                              // fast variant: the eigen iteration needs the ROBUST pass
        for (int c = tid; c < N; c += T) {
            double dP = pv.geo[G_DEG * pv.gstride + c], dN = dP;
            if (A.solve_only && A.node_wp) { dP = A.node_wp[s0 + c]; dN = A.node_wn[s0 + c]; }   // stacked-row weights of fit_horn87 / fit_irls
            pv.aP[c] = mu * dP; pv.aN[c] = dN;
        }
        __syncthreads();
        if (par.method == RP_METHOD_HORN87) {                         // rpmodule.py:60-84
            horn_fit(sh, pv, mu, red_buf);
        } else if (par.method == RP_METHOD_IRLS) {                    // rpmodule.py:169-210
            irls_rounds(sh, pv, mu, red_buf);
        } else if (par.method == RP_METHOD_IRLS_SM) {                 // rpmodule.py:212-315
            long long clk_pi = 0, clk0 = 0;
            const bool prof = A.has_dbg && A.dbg.phase_clk;
            irls_rounds(sh, pv, mu, red_buf);
            for (int alt = 0; alt < NUM_ALTER; ++alt) {
                residual_to_h(pv);
                int conv = 0;
                if (prof) clk0 = clock64();
                int it = power_iteration<false, ROBUST>(sh, pv, par.power_tol, par.max_power_iters, alt > 0, red_buf, &conv, A.pi_fast_cap, A.pi_switch);
                if (prof) { clk_pi += clock64() - clk0; if (tid == 0) A.dbg.phase_clk[(size_t)b * 8 + 6] = clk_pi; }
                tot_it += it; max_it_seen = it > max_it_seen ? it : max_it_seen; not_conv |= !conv;
                if (!ROBUST && !conv && par.max_power_iters > A.pi_fast_cap) { retry = true; break; }
                if (A.has_dbg && A.dbg.u) for (int c = tid; c < N; c += T) A.dbg.u[((size_t)b * NUM_ALTER + alt) * A.dbg.u_stride + c] = pv.ua[c];
                x_degrees(sh, pv, mu);
                irls_rounds(sh, pv, mu, red_buf);
            }
        } else if (par.method == RP_METHOD_SPECTRAL) {                // rpmodule.py:86-167
            horn_fit(sh, pv, mu, red_buf);
            for (int c = tid; c < N; c += T) pv.sv[c] = 1.0;
            __syncthreads();
            for (int alt = 0; alt < NUM_ALTER; ++alt) {
                residual_pass(sh, pv, mu, false);
                residual_to_h(pv);
                int conv = 0;
                int it = power_iteration<true, ROBUST>(sh, pv, par.power_tol, par.max_power_iters, false, red_buf, &conv, A.pi_fast_cap, A.pi_switch);
                tot_it += it; max_it_seen = it > max_it_seen ? it : max_it_seen; not_conv |= !conv;
                if (!ROBUST && !conv && par.max_power_iters > A.pi_fast_cap) { retry = true; break; }
                if (A.has_dbg && A.dbg.u) for (int c = tid; c < N; c += T) A.dbg.u[((size_t)b * NUM_ALTER + alt) * A.dbg.u_stride + c] = pv.ua[c];
                x_degrees(sh, pv, mu);
                for (int c = tid; c < N; c += T) pv.sv[c] = pv.ua[c];     // next affinity uses allWP = mu*x (:126,148)
                __syncthreads();
                horn_fit(sh, pv, 1.0, red_buf);
            }
        } else {
            write_identity(Tout);
            if (tid == 0) A.status[b] = RP_STATUS_UNSUPPORTED;
            continue;
        }
        RP_PHASE_CLK(5);
        if (retry) {                                     // uniform: decided from block-wide reductions
            if (tid == 0) A.status[b] = RP_STATUS_RETRY;
            continue;
        }
        if (tid < 16) {
            int r = tid >> 2, c = tid & 3;
            double v;
            if (r < 3) v = (c < 3) ? sh.pose.R[3 * r + c] : sh.pose.t[r];
            else v = (c == 3) ? 1.0 : 0.0;
            Tout[tid] = v;
        }
        if (tid == 0) {
            A.status[b] = RP_STATUS_OK;
            if (st) { st[4] = tot_it; st[5] = max_it_seen; st[6] = not_conv | (sh.ties << 8); if (ROBUST) st[7] |= 0x100; }
        }
    }
}

// ----------------------------------------------------------------------------------------------
struct Layout {
    size_t o_geo, o_cj, o_mask, o_edges, o_ew, o_rowstart, o_cols, o_vals, o_dyn, slot_bytes;
    int Nmax, NWmax;
    long long edge_cap;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
constexpr int TPAD = 512;

bool make_layout(int max_ns, int max_topk, long long edge_cap, size_t dyn_bytes, Layout* L) {
    long long Nmax = (long long)max_ns * max_topk;
    if (Nmax < 1 || Nmax > 16383) return false;     // 14-bit column field in the CSR word
    long long P = Nmax * (Nmax - 1) / 2;
    if (edge_cap <= 0 || edge_cap > P) edge_cap = P;
    if (edge_cap < 3) edge_cap = 3;
    L->Nmax = (int)Nmax; L->NWmax = (int)((Nmax + 31) / 32); L->edge_cap = edge_cap;
    size_t o = 0;
    L->o_geo = o; o = align_up(o + sizeof(double) * G_COUNT * Nmax, 256);
    L->o_cj = o; o = align_up(o + sizeof(int) * Nmax, 256);
    L->o_mask = o; o = align_up(o + sizeof(unsigned) * Nmax * L->NWmax, 256);
    L->o_edges = o; o = align_up(o + sizeof(unsigned) * edge_cap, 256);
    L->o_ew = o; o = align_up(o + sizeof(double) * edge_cap, 256);
    L->o_rowstart = o; o = align_up(o + sizeof(uint16_t) * Nmax * L->NWmax, 256);
    L->o_cols = o; o = align_up(o + sizeof(uint16_t) * (2 * edge_cap + 2 * TPAD), 256);  // lane-interleaved: up to T-1 pad (TPAD: the
                                                                                          // widest build, so both builds share one layout)
    L->o_vals = o; o = align_up(o + sizeof(double) * (2 * edge_cap + 2 * TPAD), 256);
    L->o_dyn = o; o = align_up(o + dyn_bytes, 256);          // only for pairs whose vectors do not fit shared memory
    L->slot_bytes = o;
    return true;
}

int default_slots_uncached(size_t smem_bytes);

int pi_switch() {
    static int v = 0;
    if (v == 0) { const char* e = getenv("RP_PI_SWITCH"); v = e ? atoi(e) : PI_SWITCH; if (v < 4) v = 4; }
    return v;
}

int pi_fast_cap() {                 // PI_FAST_CAP, or RP_PI_FAST_CAP from the environment (tuning aid)
    static int v = 0;
    if (v == 0) { const char* e = getenv("RP_PI_FAST_CAP"); v = e ? atoi(e) : PI_FAST_CAP; if (v < 8) v = 8; }
    return v;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize of both kernel variants, raised on demand and never lowered (launches with
// different plans may be in flight on several streams; the attribute only has to cover the largest of them)
template <bool ROBUST, bool BIG> __global__ void rp_solve_kernel(const SolveArgs A);
bool ensure_smem_attr(size_t bytes) {
    static size_t have_dev[64];                         // per device: function attributes are per-device state
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); dev = 63; have_dev[63] = 0; }
    size_t& have = have_dev[dev];
    if (have == 0) have = 48 * 1024;
    if (bytes <= have) return true;
    if (cudaFuncSetAttribute(rp_solve_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess ||
        cudaFuncSetAttribute(rp_solve_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    have = bytes;
    return true;
}

struct SmemPlan { size_t bytes; int mask_in_smem; int tfeat_stride; size_t mask_off; int dyn_in_global; size_t dyn_bytes; };

bool make_smem_plan(long long Nmax_, int max_nt, int feat_dim, SmemPlan* S) {
    if (Nmax_ < 1 || Nmax_ > 16383) return false;
    struct { int Nmax, NWmax; } L = {(int)Nmax_, (int)((Nmax_ + 31) / 32)};
    // ~227 KB per SM shared by RP_MIN_BLOCKS CTAs; the static part (struct Shared) is ~3-8 KB
    const size_t budget = (size_t)(220 * 1024) / RP_MIN_BLOCKS - 9 * 1024;
    size_t hard = (size_t)227 * 1024 - sizeof(Shared) - 2048;     // dynamic + static shared memory of one CTA <= 227 KB
    if (hard > (size_t)200 * 1024) hard = (size_t)200 * 1024;
    int ts = (feat_dim % 8 == 0) ? (feat_dim + 4) : (feat_dim | 1);   // 16-B aligned rows, (ts/4) odd -> conflict-free LDS.128
    size_t fe = (size_t)max_nt * ts * sizeof(float);
    size_t vec = align_up((size_t)8 * L.Nmax * sizeof(double) + (size_t)(2 * L.Nmax + 1) * sizeof(int), 16);   // 6 vectors + S + rowstart + rowmap
    size_t mask = (size_t)L.Nmax * L.NWmax * sizeof(unsigned);
    size_t base = fe > vec ? fe : vec;
    base = align_up(base, 16);
    int in_smem = (base + mask <= budget) ? 1 : 0;
    size_t need = base + (in_smem ? mask : 0);
    S->tfeat_stride = ts; S->mask_off = base;
    if (need > hard) {          // big pair (n_s * topK > ~2800 or n_t > ~1400): same layout in the slot's global workspace
        S->bytes = 0; S->mask_in_smem = 0; S->dyn_in_global = 1; S->dyn_bytes = base;
        return true;
    }
    S->bytes = need; S->mask_in_smem = in_smem; S->dyn_in_global = 0; S->dyn_bytes = 0;
    return true;
}

int sm_count() {
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm < 1) {
            cudaGetLastError();
            n_sm = 148;
        }
    }
    return n_sm;
}

// Launch-time shared-memory plan.  With fewer CTAs than an SM can hold (small batches: B < 4 x 148) the idle shared memory
// goes to the resident pairs: first the symmetric bit mask, then as much of the CSR of W as fits (cols uint16 + vals
// float64, 10 bytes per directed non-zero).  The fitters make ~250 passes over that CSR per pair; from shared memory a pass
// costs a few microseconds instead of an L2 round trip per batch of loads (the L1 of a CTA with a 200 KB carve-out is tiny).
struct LaunchPlan { size_t bytes; int mask_in_smem; int csr_cap; size_t csr_off; size_t geo_off; };
LaunchPlan make_launch_plan(const SmemPlan& S, const Layout& L, int grid) {
    LaunchPlan P = {S.bytes, S.mask_in_smem, 0, 0, 0};
    if (S.dyn_in_global) return P;
    const int per_sm = (grid + sm_count() - 1) / sm_count();
    if (RP_MIN_BLOCKS > 1 && per_sm >= RP_MIN_BLOCKS) return P;
    const size_t stat = sizeof(Shared) + 2560;                                  // static part + 1 KB reserved per CTA + slack
    size_t avail = (size_t)(227 * 1024) / (size_t)per_sm - stat;
    if (avail > (size_t)225 * 1024 - stat) avail = (size_t)225 * 1024 - stat;
    // the CSR first (read ~250 times per pair), the bit mask (written in D, read once in E) only if both fit
    const size_t mask = (size_t)L.Nmax * L.NWmax * sizeof(unsigned);
    const long long want = (2 * L.edge_cap + 2 * T + 7) & ~7ll;
    if (!P.mask_in_smem && align_up(S.mask_off + mask, 16) + (size_t)want * 10 <= avail) { P.mask_in_smem = 1; P.bytes = S.mask_off + mask; }
    if (RP_MIN_BLOCKS == 1) {                              // one CTA per SM: room for the fitters' geometry (12 N float64)
        const size_t g0 = align_up(P.bytes, 16), gb = (size_t)12 * L.Nmax * sizeof(double);
        if (g0 > 0 && g0 + gb + 8192 <= avail) { P.geo_off = g0; P.bytes = g0 + gb; }
    }
    const size_t cur = align_up(P.bytes, 16);
    if (avail > cur + 4096) {
        long long entries = (long long)((avail - cur) / 10) & ~7ll;               // vals start 16-byte aligned
        if (entries > want) entries = want;
        P.csr_cap = (int)entries; P.csr_off = cur; P.bytes = cur + (size_t)entries * 10;
    }
    return P;
}

int default_slots(size_t smem_bytes) {
    static size_t cached_bytes = (size_t)-1;
    static int cached_slots = 0;
    if (smem_bytes == cached_bytes && cached_slots > 0) return cached_slots;
    int slots = default_slots_uncached(smem_bytes);
    if (slots > 0) { cached_bytes = smem_bytes; cached_slots = slots; }
    return slots;
}

int default_slots_uncached(size_t smem_bytes) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    int per = 0;
    if (!ensure_smem_attr(smem_bytes)) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, rp_solve_kernel<false, false>, T, smem_bytes) != cudaSuccess) return -1;
    if (per < 1) per = 1;
    return sms * per;
}


// Two builds of this file make up the solver: the default one (T = 128, four CTAs per SM: throughput for large batches, and
// every extern "C" entry point) and rp_solver_wide.cu (RP_WIDE_TU: T = 512, one CTA per SM), whose two launch functions below
// take small batches -- with fewer scan pairs than SMs x 2 a pair gets a whole SM's threads instead of a quarter of them.  Same
// kernel source, same workspace layout; reduction trees differ with T, so the two builds agree to rounding (1e-13 on the
// poses), not bitwise: a batch goes through exactly one of them, chosen from its size alone (rp_solver_wide_max).
constexpr int RP_WIDE_DECLINED = 1000;       // the wide build cannot take this shape (vectors beyond shared memory): use T = 128
int solve_batch_impl(int B, const int32_t* off_s, const int32_t* off_t,
                      const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                      const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                      int feat_dim, const rp_params* params, const int32_t* param_idx,
                      const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                      int max_ns, int max_nt, int max_topk,
                      int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                      double* T_out, int32_t* status, int32_t* stats,
                      int stop_after, const rp_debug* dbg, void* stream_) {
    if (B < 0 || !off_s || !off_t || !feat_s || !feat_t || !w_s || !w_t || !params || !workspace || !status)
        return RP_ERR_INVALID_ARG;
    if (stop_after == RP_STAGE_SOLVE && !T_out) return RP_ERR_INVALID_ARG;
    if (stop_after != RP_STAGE_TOPK && (!pc_s || !pc_t || !nrm_s || !nrm_t)) return RP_ERR_INVALID_ARG;
    if (stop_after < RP_STAGE_TOPK || stop_after > RP_STAGE_SOLVE) return RP_ERR_INVALID_ARG;
    if (max_topk > RP_MAX_TOPK || max_topk < 1 || feat_dim < 1 || feat_dim > RP_MAX_FEAT_DIM) return RP_ERR_UNSUPPORTED;
    if (B == 0) return RP_OK;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    SmemPlan S;
    if (!make_smem_plan((long long)max_ns * max_topk, max_nt, feat_dim, &S)) return RP_ERR_UNSUPPORTED;
    Layout L;
    if (!make_layout(max_ns, max_topk, edge_cap, S.dyn_bytes, &L)) return RP_ERR_UNSUPPORTED;
#ifdef RP_WIDE_TU
    if (S.dyn_in_global) return RP_WIDE_DECLINED;
    {
        const int mine = default_slots(S.bytes);
        if (mine < 0) { cudaGetLastError(); return RP_ERR_NO_DEVICE; }
        if (n_slots <= 0 || n_slots > mine) n_slots = mine;
    }
#else
    if (n_slots <= 0) {
        n_slots = default_slots(S.bytes);
        if (n_slots < 0) { cudaGetLastError(); return RP_ERR_NO_DEVICE; }
    }
#endif
    const int grid = B < n_slots ? B : n_slots;
    const LaunchPlan P = make_launch_plan(S, L, grid);
    if (!ensure_smem_attr(P.bytes)) return RP_ERR_CUDA;
    if (workspace_bytes < 256 + (size_t)n_slots * L.slot_bytes) return RP_ERR_WORKSPACE_TOO_SMALL;
    SolveArgs a;
    a.B = B; a.off_s = off_s; a.off_t = off_t;
    a.pc_s = pc_s; a.nrm_s = nrm_s; a.feat_s = feat_s; a.w_s = w_s;
    a.pc_t = pc_t; a.nrm_t = nrm_t; a.feat_t = feat_t; a.w_t = w_t;
    a.feat_dim = feat_dim; a.params = params; a.param_idx = param_idx; a.zero_row_topk = zero_row_topk; a.feat_sum_order = feat_sum_order;
    a.solve_only = 0; a.node_wp = nullptr; a.node_wn = nullptr; a.edge_off = nullptr; a.edge_rc = nullptr; a.edge_w = nullptr;
    a.max_topk = max_topk; a.edge_cap = L.edge_cap;
    a.ws = static_cast<char*>(workspace); a.slot_bytes = L.slot_bytes;
    a.o_geo = L.o_geo; a.o_cj = L.o_cj; a.o_mask = L.o_mask; a.o_edges = L.o_edges; a.o_ew = L.o_ew;
    a.o_rowstart = L.o_rowstart; a.o_cols = L.o_cols; a.o_vals = L.o_vals;
    a.Nmax = L.Nmax; a.NWmax = L.NWmax;
    a.T_out = T_out; a.status = status; a.stats = stats;
    a.stop_after = stop_after;
    a.has_dbg = dbg ? 1 : 0;
    if (dbg) a.dbg = *dbg; else { rp_debug z = {}; a.dbg = z; }
    a.mask_in_smem = P.mask_in_smem; a.tfeat_stride = S.tfeat_stride; a.sm_mask_off = S.mask_off;
    a.csr_smem_cap = P.csr_cap; a.sm_csr_off = P.csr_off; a.sm_geo_off = P.geo_off; a.pi_fast_cap = pi_fast_cap(); a.pi_switch = pi_switch();
    a.dyn_in_global = S.dyn_in_global; a.o_dyn = L.o_dyn;
    if (cudaMemsetAsync(workspace, 0, 256, stream) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
    if (S.dyn_in_global) rp_solve_kernel<false, true><<<grid, T, P.bytes, stream>>>(a);
    else rp_solve_kernel<false, false><<<grid, T, P.bytes, stream>>>(a);
    ++g_launches;
    if (a.stop_after == RP_STAGE_SOLVE) {                // pairs whose eigen iteration needs the accelerated method
        if (S.dyn_in_global) rp_solve_kernel<true, true><<<grid, T, P.bytes, stream>>>(a);
        else rp_solve_kernel<true, false><<<grid, T, P.bytes, stream>>>(a);
        ++g_launches;
    }
    if (cudaGetLastError() != cudaSuccess) return RP_ERR_CUDA;
    return RP_OK;
}


int spectral_irls_impl(int B, const int32_t* node_off,
                           const double* sp, const double* sn, const double* tp, const double* tn,
                           const double* node_wp, const double* node_wn,
                           const int32_t* edge_off, const int32_t* edge_rc, const double* edge_w,
                           const rp_params* params, const int32_t* param_idx, int max_nodes,
                           int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                           double* T_out, int32_t* status, int32_t* stats, void* stream_) {
    if (B < 0 || !node_off || !sp || !sn || !tp || !tn || !params || !workspace || !T_out || !status) return RP_ERR_INVALID_ARG;
    if (!edge_off && !(node_wp && node_wn)) return RP_ERR_INVALID_ARG;
    if (edge_off && (!edge_rc || !edge_w)) return RP_ERR_INVALID_ARG;
    if (B == 0) return RP_OK;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    SmemPlan S;
    if (!make_smem_plan(max_nodes, 1, 8, &S)) return RP_ERR_UNSUPPORTED;
    Layout L;
    if (!make_layout(max_nodes, 1, edge_cap, S.dyn_bytes, &L)) return RP_ERR_UNSUPPORTED;
#ifdef RP_WIDE_TU
    if (S.dyn_in_global) return RP_WIDE_DECLINED;
    {
        const int mine = default_slots(S.bytes);
        if (mine < 0) { cudaGetLastError(); return RP_ERR_NO_DEVICE; }
        if (n_slots <= 0 || n_slots > mine) n_slots = mine;
    }
#else
    if (n_slots <= 0) {
        n_slots = default_slots(S.bytes);
        if (n_slots < 0) { cudaGetLastError(); return RP_ERR_NO_DEVICE; }
    }
#endif
    const int grid = B < n_slots ? B : n_slots;
    const LaunchPlan P = make_launch_plan(S, L, grid);
    if (!ensure_smem_attr(P.bytes)) return RP_ERR_CUDA;
    if (workspace_bytes < 256 + (size_t)n_slots * L.slot_bytes) return RP_ERR_WORKSPACE_TOO_SMALL;
    SolveArgs a = {};
    a.B = B; a.off_s = node_off; a.off_t = node_off;
    a.pc_s = sp; a.nrm_s = sn; a.pc_t = tp; a.nrm_t = tn;
    a.feat_dim = 8; a.params = params; a.param_idx = param_idx;
    a.solve_only = 1; a.node_wp = node_wp; a.node_wn = node_wn; a.edge_off = edge_off; a.edge_rc = edge_rc; a.edge_w = edge_w;
    a.max_topk = 1; a.edge_cap = L.edge_cap;
    a.ws = static_cast<char*>(workspace); a.slot_bytes = L.slot_bytes;
    a.o_geo = L.o_geo; a.o_cj = L.o_cj; a.o_mask = L.o_mask; a.o_edges = L.o_edges; a.o_ew = L.o_ew;
    a.o_rowstart = L.o_rowstart; a.o_cols = L.o_cols; a.o_vals = L.o_vals;
    a.Nmax = L.Nmax; a.NWmax = L.NWmax;
    a.T_out = T_out; a.status = status; a.stats = stats;
    a.stop_after = RP_STAGE_SOLVE; a.has_dbg = 0;
    a.mask_in_smem = P.mask_in_smem; a.tfeat_stride = S.tfeat_stride; a.sm_mask_off = S.mask_off;
    a.csr_smem_cap = P.csr_cap; a.sm_csr_off = P.csr_off; a.sm_geo_off = P.geo_off; a.pi_fast_cap = pi_fast_cap(); a.pi_switch = pi_switch();
    a.dyn_in_global = S.dyn_in_global; a.o_dyn = L.o_dyn;
    if (cudaMemsetAsync(workspace, 0, 256, stream) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
    if (S.dyn_in_global) rp_solve_kernel<false, true><<<grid, T, P.bytes, stream>>>(a);
    else rp_solve_kernel<false, false><<<grid, T, P.bytes, stream>>>(a);
    ++g_launches;
    if (a.stop_after == RP_STAGE_SOLVE) {                // pairs whose eigen iteration needs the accelerated method
        if (S.dyn_in_global) rp_solve_kernel<true, true><<<grid, T, P.bytes, stream>>>(a);
        else rp_solve_kernel<true, false><<<grid, T, P.bytes, stream>>>(a);
        ++g_launches;
    }
    if (cudaGetLastError() != cudaSuccess) return RP_ERR_CUDA;
    return RP_OK;
}


}  // namespace

// launch functions of the T = 512 build (defined by rp_solver_wide.cu, called from the entry points below)
int rp_wide_solve_batch_ex(int B, const int32_t* off_s, const int32_t* off_t,
                      const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                      const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                      int feat_dim, const rp_params* params, const int32_t* param_idx,
                      const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                      int max_ns, int max_nt, int max_topk,
                      int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                      double* T_out, int32_t* status, int32_t* stats,
                      int stop_after, const rp_debug* dbg, void* stream_);
int rp_wide_spectral_irls_solve(int B, const int32_t* node_off,
                           const double* sp, const double* sn, const double* tp, const double* tn,
                           const double* node_wp, const double* node_wn,
                           const int32_t* edge_off, const int32_t* edge_rc, const double* edge_w,
                           const rp_params* params, const int32_t* param_idx, int max_nodes,
                           int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                           double* T_out, int32_t* status, int32_t* stats, void* stream_);
long long rp_wide_launch_count();

#ifdef RP_WIDE_TU
int rp_wide_solve_batch_ex(int B, const int32_t* off_s, const int32_t* off_t,
                      const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                      const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                      int feat_dim, const rp_params* params, const int32_t* param_idx,
                      const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                      int max_ns, int max_nt, int max_topk,
                      int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                      double* T_out, int32_t* status, int32_t* stats,
                      int stop_after, const rp_debug* dbg, void* stream_) {
    return solve_batch_impl(B, off_s, off_t, pc_s, nrm_s, feat_s, w_s, pc_t, nrm_t, feat_t, w_t, feat_dim, params, param_idx, zero_row_topk, feat_sum_order, max_ns, max_nt, max_topk, n_slots, edge_cap, workspace, workspace_bytes, T_out, status, stats, stop_after, dbg, stream_);
}
int rp_wide_spectral_irls_solve(int B, const int32_t* node_off,
                           const double* sp, const double* sn, const double* tp, const double* tn,
                           const double* node_wp, const double* node_wn,
                           const int32_t* edge_off, const int32_t* edge_rc, const double* edge_w,
                           const rp_params* params, const int32_t* param_idx, int max_nodes,
                           int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                           double* T_out, int32_t* status, int32_t* stats, void* stream_) {
    return spectral_irls_impl(B, node_off, sp, sn, tp, tn, node_wp, node_wn, edge_off, edge_rc, edge_w, params, param_idx, max_nodes, n_slots, edge_cap, workspace, workspace_bytes, T_out, status, stats, stream_);
}
long long rp_wide_launch_count() { return g_launches; }
#else

namespace {
// Batches of at most this many scan pairs run on the T = 512 build (0: never).  Default = the SM count: one wave of whole-SM
// pairs (measured, N = 515: 0.42 vs 0.53 ms up to 148 pairs; at 222 pairs the second wave makes it 0.82 vs 0.60 ms).
// RP_SOLVER_WIDE_MAX in the environment or rp_solver_wide_max() override it.
int g_wide_max = -1;
int wide_max() {
    if (g_wide_max < 0) {
        const char* e = getenv("RP_SOLVER_WIDE_MAX");
        g_wide_max = e ? atoi(e) : sm_count();
        if (g_wide_max < 0) g_wide_max = 0;
    }
    return g_wide_max;
}
}  // namespace

extern "C" {

int rp_abi_version(void) { return RP_ABI_VERSION; }

int64_t rp_launch_count(void) { return g_launches + rp_wide_launch_count(); }

int rp_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return RP_ERR_NO_DEVICE; }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { cudaGetLastError(); return RP_ERR_NO_DEVICE; }
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return RP_OK;
}

int rp_solve_workspace_bytes(int n_slots, int max_ns, int max_nt, int max_topk, int feat_dim,
                             int64_t edge_cap, size_t* bytes) {
    if (!bytes || max_ns < 1 || max_nt < 1 || max_topk < 1) return RP_ERR_INVALID_ARG;
    if (max_topk > RP_MAX_TOPK || feat_dim < 1 || feat_dim > RP_MAX_FEAT_DIM) return RP_ERR_UNSUPPORTED;
    SmemPlan S;
    if (!make_smem_plan((long long)max_ns * max_topk, max_nt, feat_dim, &S)) return RP_ERR_UNSUPPORTED;
    Layout L;
    if (!make_layout(max_ns, max_topk, edge_cap, S.dyn_bytes, &L)) return RP_ERR_UNSUPPORTED;
    if (n_slots <= 0) {
        n_slots = default_slots(S.bytes);
        if (n_slots < 0) { cudaGetLastError(); return RP_ERR_NO_DEVICE; }
    }
    *bytes = 256 + (size_t)n_slots * L.slot_bytes;
    return RP_OK;
}

int rp_solve_default_slots(int max_ns, int max_nt, int max_topk, int feat_dim, int* n_slots) {
    if (!n_slots || max_ns < 1 || max_nt < 1 || max_topk < 1) return RP_ERR_INVALID_ARG;
    if (max_topk > RP_MAX_TOPK || feat_dim < 1 || feat_dim > RP_MAX_FEAT_DIM) return RP_ERR_UNSUPPORTED;
    SmemPlan S;
    if (!make_smem_plan((long long)max_ns * max_topk, max_nt, feat_dim, &S)) return RP_ERR_UNSUPPORTED;
    int n = default_slots(S.bytes);
    if (n < 0) { cudaGetLastError(); return RP_ERR_NO_DEVICE; }
    *n_slots = n;
    return RP_OK;
}

int rp_solve_batch_ex(int B, const int32_t* off_s, const int32_t* off_t,
                      const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                      const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                      int feat_dim, const rp_params* params, const int32_t* param_idx,
                      const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                      int max_ns, int max_nt, int max_topk,
                      int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                      double* T_out, int32_t* status, int32_t* stats,
                      int stop_after, const rp_debug* dbg, void* stream_) {
    if (B > 0 && B <= wide_max()) {
        const int rc = rp_wide_solve_batch_ex(B, off_s, off_t, pc_s, nrm_s, feat_s, w_s, pc_t, nrm_t, feat_t, w_t, feat_dim, params, param_idx, zero_row_topk, feat_sum_order, max_ns, max_nt, max_topk, n_slots, edge_cap, workspace, workspace_bytes, T_out, status, stats, stop_after, dbg, stream_);
        if (rc != RP_WIDE_DECLINED) return rc;
    }
    return solve_batch_impl(B, off_s, off_t, pc_s, nrm_s, feat_s, w_s, pc_t, nrm_t, feat_t, w_t, feat_dim, params, param_idx, zero_row_topk, feat_sum_order, max_ns, max_nt, max_topk, n_slots, edge_cap, workspace, workspace_bytes, T_out, status, stats, stop_after, dbg, stream_);
}

int rp_solver_wide_max(int new_max) {
    const int old = wide_max();
    if (new_max >= 0) g_wide_max = new_max;
    return old;
}

int rp_solve_batch(int B, const int32_t* off_s, const int32_t* off_t,
                   const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                   const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                   int feat_dim, const rp_params* params, const int32_t* param_idx,
                   const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                   int max_ns, int max_nt, int max_topk,
                   int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                   double* T_out, int32_t* status, int32_t* stats, void* stream) {
    return rp_solve_batch_ex(B, off_s, off_t, pc_s, nrm_s, feat_s, w_s, pc_t, nrm_t, feat_t, w_t, feat_dim, params,
                             param_idx, zero_row_topk, feat_sum_order, max_ns, max_nt, max_topk, n_slots, edge_cap, workspace,
                             workspace_bytes, T_out, status, stats, RP_STAGE_SOLVE, nullptr, stream);
}

int rp_match_topk(int B, const int32_t* off_s, const int32_t* off_t,
                  const float* feat_s, const double* w_s, const float* feat_t, const double* w_t,
                  int feat_dim, const rp_params* params, const int32_t* param_idx,
                  const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                  int max_ns, int max_nt, int max_topk,
                  int n_slots, void* workspace, size_t workspace_bytes,
                  int32_t* topk_idx, double* topk_f, int32_t* status, void* stream) {
    if (!topk_idx || !status) return RP_ERR_INVALID_ARG;
    rp_debug d = {};
    d.topk_idx = topk_idx; d.topk_f = topk_f;
    return rp_solve_batch_ex(B, off_s, off_t, nullptr, nullptr, feat_s, w_s, nullptr, nullptr, feat_t, w_t, feat_dim,
                             params, param_idx, zero_row_topk, feat_sum_order, max_ns, max_nt, max_topk, n_slots, 0, workspace,
                             workspace_bytes, nullptr, status, nullptr, RP_STAGE_TOPK, &d, stream);
}

// Stage entry rpmodule.py:342-472: the surviving (geometrically consistent) pairs and their weights.
int rp_affinity_build(int B, const int32_t* off_s, const int32_t* off_t,
                      const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                      const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                      int feat_dim, const rp_params* params, const int32_t* param_idx,
                      const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                      int max_ns, int max_nt, int max_topk,
                      int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                      int32_t* topk_idx, int32_t* edge_rc, double* edge_w, int32_t* status, int32_t* stats, void* stream) {
    if (!edge_rc || !edge_w || !status || edge_cap < 1) return RP_ERR_INVALID_ARG;
    rp_debug d = {};
    d.topk_idx = topk_idx; d.edge_rc = edge_rc; d.edge_w = edge_w; d.edge_cap = edge_cap;
    return rp_solve_batch_ex(B, off_s, off_t, pc_s, nrm_s, feat_s, w_s, pc_t, nrm_t, feat_t, w_t, feat_dim,
                             params, param_idx, zero_row_topk, feat_sum_order, max_ns, max_nt, max_topk, n_slots, 0, workspace,
                             workspace_bytes, nullptr, status, stats, RP_STAGE_AFFINITY, &d, stream);
}

// Stage entry rpmodule.py:484-508 (fitters only): B problems, each a set of correspondences ("nodes": source/target
// position and normal) and an optional list of consistent pairs ("edges": two node indices and the pair weight w,
// rpmodule.py:457-467).  With edges: base weights are the row degrees of W (fit_spectral / fit_irls_sm; also
// fit_horn87 / fit_irls on the helper's stacked rows).  Without edges: explicit per-node weights (allWP, allWN of
// fit_horn87 / fit_irls; horn87_np = normals only).  Replaces fit_horn87 :60, fit_spectral :86, fit_irls :169,
// fit_irls_sm :212 and horn87_np :17.
int rp_spectral_irls_solve(int B, const int32_t* node_off,
                           const double* sp, const double* sn, const double* tp, const double* tn,
                           const double* node_wp, const double* node_wn,
                           const int32_t* edge_off, const int32_t* edge_rc, const double* edge_w,
                           const rp_params* params, const int32_t* param_idx, int max_nodes,
                           int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                           double* T_out, int32_t* status, int32_t* stats, void* stream_) {
    if (B > 0 && B <= wide_max()) {
        const int rc = rp_wide_spectral_irls_solve(B, node_off, sp, sn, tp, tn, node_wp, node_wn, edge_off, edge_rc, edge_w, params, param_idx, max_nodes, n_slots, edge_cap, workspace, workspace_bytes, T_out, status, stats, stream_);
        if (rc != RP_WIDE_DECLINED) return rc;
    }
    return spectral_irls_impl(B, node_off, sp, sn, tp, tn, node_wp, node_wn, edge_off, edge_rc, edge_w, params, param_idx, max_nodes, n_slots, edge_cap, workspace, workspace_bytes, T_out, status, stats, stream_);
}

int rp_spectral_irls_workspace_bytes(int n_slots, int max_nodes, int64_t edge_cap, size_t* bytes) {
    return rp_solve_workspace_bytes(n_slots, max_nodes, 1, 1, 8, edge_cap, bytes);
}

}  // extern "C"
#endif  // RP_WIDE_TU

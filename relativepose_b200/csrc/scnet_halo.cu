// scnet_halo.cu -- halo-tile implicit-GEMM convolution / transposed convolution on tcgen05 (sm_100a only).
//
// scnet_tc.cu gathers one [128 pixel x TK channel] A tile per (tap, K chunk): a 4x4 kernel re-reads (and re-normalises,
// re-converts) every input pixel 16 times, so those layers are gather-bound (tensor pipe ~5 %).  Here a CTA owns a
// 16 x 8 block of output positions and stages the INPUT HALO of that block once per K chunk:
//
//   smem halo  [K core kc][halo pixel h] 16-byte units (8 bf16 channels), h = plane, row, column
//
// In the canonical K-major / no-swizzle UMMA layout a core matrix is 8 rows x 16 bytes; 8 horizontally adjacent
// halo pixels ARE one core matrix, the next 8 rows of the M=128 tile are the next halo row (SBO = row pitch x 16 B)
// and the next 8 channels are the next K core (LBO).  The A operand of tap (dy,dx) is therefore the same halo with
// the descriptor start address moved by (dy x pitch + dx) x 16 bytes -- no data movement per tap at all.
//   * stride-2 convolutions split the halo into 4 parity planes (iy = 2(a+e)+q), so rows stay 16 bytes apart;
//   * a stride-2 transposed convolution is its 4 sub-pixel classes over ONE halo: 16 (class, tap) MMAs per K step
//     into 4 accumulators in TMEM (4 x BN columns).
// Persistent and warp-specialised (one CTA per SM looping over tiles): warps 4-11 gather (producer BatchNorm +
// LeakyReLU + bf16 convert, 16-byte STS; 2 or 4 loader groups, each filling its own halo buffer), warp 12 streams the
// pre-packed weight blocks with cp.async.bulk (1-D TMA) through an NB-deep ring (or loads them once when they all fit),
// warp 13 issues tcgen05.mma into one of 2-8 accumulator sets in TMEM, warps 0-3 run the epilogue of earlier tiles
// meanwhile (tcgen05.ld, bias/tanh, raw output fp32 or bf16 NHWC, deterministic per-tile partial batch statistics).
// mbarriers only; no __syncthreads in the steady state.
#include "rp_h16.cuh"
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/rp_b200.h"
#include "scnet_common.cuh"
#include "tc_prims.cuh"

namespace halo {
using namespace tc;

constexpr int TH = 16, TW = 8, TM = 128;
constexpr int EPI = 128, LOADERS = 256, CTA = EPI + LOADERS + 64;    // warps 0-3 epilogue, 4-11 loaders, 12 TMA, 13 MMA
constexpr int MAXT = 16;            // taps over all classes
constexpr int MAXACC = 8;           // accumulator sets in TMEM (512 columns / (classes x BN), at most 8)
constexpr int MAXNPX = 640;         // halo pixels (4 planes x 17 x 9 = 612 for 4x4 s2)
constexpr int MAXNG = 8;            // loader groups = halo buffers in flight (2, 4 or 8)
constexpr int PIXTAB = 2 * MAXNPX;  // ints of per-group pixel tables
constexpr int EPI_SMEM = (4 * 32 * 33 + 16 * 4 * 64) * 4;            // per-warp transpose buffers + partial sums

// Profiling only (flags bit 5): cycles per role, summed over CTAs -- [0] loader total, [1] loader wait a_empty, [2] loader copy wait,
// [3] loader transform, [4] mma total, [5] mma wait acc_empty, [6] mma wait a_full, [7] mma wait w_full, [8] epilogue total,
// [9] epilogue wait acc_full, [10] epilogue tmem_ld, [11] epilogue stats barriers, [12] loader table build,
// [13] loader copy issue, [14] loader fence + arrive
__device__ unsigned long long g_halo_prof[16];
#define HP_T0(var) long long var = 0; if (PROF) var = clock64();
#define HP_ADD(idx, var) if (PROF) { prof_acc[idx] += clock64() - var; }

struct Plane { int qy, qx, oy, ox; };      // input row = (a0 + oy + hy) * istr + qy (same for columns)
struct HTap { int cls, a_off, first; };    // a_off: byte offset of the tap's first row inside one K core of the halo
struct HaloArgs {
    rp_conv_src src[2];
    int nsrc;
    int G, gsz, Hin, Win, Hout, Wout, Cout;
    int istr, ostr, nclass;
    int cls_py[4], cls_px[4], cls_Ha[4], cls_Wb[4];
    int nty, ntx, tiles_m, ntn, total_tiles;
    float inv_ntn, inv_tiles_m, inv_tiles_img, inv_ntx;   // reciprocals for fdiv()
    int nplane, PH, PW, NPX;               // uniform plane size, PW = row pitch in pixels
    int a_lbo;                             // bytes between the K cores of the halo (16 mod 128: conflict-free STS.128)
    Plane plane[4];
    int ntap;
    HTap tap[MAXT];
    int nkt, nkt0;                         // K chunks in total / in source 0
    int acc_cols, tmem_cols, nacc;         // TMEM columns of one accumulator set (nclass x BN) / allocated / number of sets
    int ng;                                // loader groups = halo buffers (2 or 4): group i fills buffer i with chunks i, i+ng, ...
    float inv_nkt;
    int w_resident;                        // all weight blocks of a tile fit in the ring: loaded once per CTA, never freed
    float inv_ppl;                         // 1 / (PH * PW)
    int use_tma;                           // 16-bit sources: the raw halo arrives by tiled TMA (one 5-D box per parity plane and K chunk)
    int plane_bytes, a_bytes;              // TMA layout: [plane][K core][row][pixel] 16-byte units, planes 128-byte aligned; bytes of one buffer
    int dbg;                               // profiling only (flags >> 2): 1 no epilogue stores/stats, 2 no loader copy/transform, 4 no MMA
    int h2math;                            // producer BatchNorm + LeakyReLU of 16-bit sources in packed half arithmetic (flags bit 1)
    // Split-precision launches (SPLIT instantiation; float32 sources and output; scnet_engine.py mode 'tc3'):
    //   x w = half(x') hi(w') + lo(x') hi(w') + half(x') lo(w'),  x' = 2^4 x, w' = 2^8 w, lo(v) = v - half(v)  (both low halves are
    //   normal IEEE-half numbers for |x| > 0.008), out_scale = 2^-12 applied in the epilogue.
    int split_lo;                          // flags bit 8: the loader emits lo(x') instead of half(x')
    int accum;                             // flags bit 9: out = out + out_scale * (this launch's accumulators)
    int fused3;                            // flags bit 10: ONE launch, all three terms: the loader fills a hi and a lo halo (a_half bytes apart), the weight
    int a_half;                            //   blocks come as (hi(w'), lo(w')) pairs, three MMAs per tap and K step into one accumulator
    int pairw;                             // flags bit 11: single halo, (hi, lo) weight block pairs, two MMAs per step: half(x') [hi(w') + lo(w')] (or, with
                                           //   bit 8, lo(x') ...) -- the first of the TWO launches of layers whose doubled halo does not fit shared
                                           //   memory (stride-2 convolutions); the second is bit 8 + bit 9 with the hi(w') blocks alone
    float in_scale, out_scale;
    void* out; int out_pitch, out_ch_off, out_bf16;
    float* psum; float* psq;
    const float* bias; int tanh_out;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// n / d for 0 <= n < 2^24 through a float reciprocal and one correction step (the integer divide sequence costs ~20
// instructions; a persistent CTA decodes a tile index per role and per tile)
__device__ __forceinline__ int fdiv(int n, int d, float inv) {
    int q = (int)((float)n * inv);
    if (q * d > n) --q;
    if ((q + 1) * d <= n) ++q;
    return q;
}

struct TileCoord { int g, tile_m, tile_n, img, a0, b0; };
__device__ __forceinline__ TileCoord decode_tile(const HaloArgs& A, int t) {
    TileCoord c;
    const int r = fdiv(t, A.ntn, A.inv_ntn);
    c.tile_n = t - r * A.ntn;
    c.g = fdiv(r, A.tiles_m, A.inv_tiles_m);
    c.tile_m = r - c.g * A.tiles_m;
    const int tiles_img = A.nty * A.ntx;
    const int im = fdiv(c.tile_m, tiles_img, A.inv_tiles_img);
    const int trem = c.tile_m - im * tiles_img;
    const int tyi = fdiv(trem, A.ntx, A.inv_ntx);
    c.a0 = tyi * TH; c.b0 = (trem - tyi * A.ntx) * TW;
    c.img = c.g * A.gsz + im;
    return c;
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint64_t pack_desc(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; }

// Persistent: CTA b works on tiles b, b + gridDim.x, ...  (tile = scan pair, 16x8 block of output positions, n-tile).
// Four roles run the same tile sequence and meet only at mbarriers:
//   loaders  (256 thr)  halo of K chunk c -> smem buffer c % ng (group c % ng)     a_empty -> a_full
//   TMA      (1 thr)    weight block (chunk, tap) -> ring slot (or all blocks once)  w_empty -> w_full
//   MMA      (1 thr)    ntap x TK/16 tcgen05.mma per chunk into accumulator set tile % nacc; commits free slots / buffers
//   epilogue (128 thr)  accumulator set tile % nacc -> global + statistics          acc_full -> acc_empty
// so the gathers of the next tiles, the MMAs of tile i and the epilogue of earlier tiles overlap.
template <int BN, int TK, int NB, bool SPLIT>      // SPLIT: the launches of the split-precision mode (HaloArgs::split_lo / accum)
__global__ void __launch_bounds__(CTA, 1) conv_halo_tc(const HaloArgs A, const unsigned char* __restrict__ Wp,
                                                        const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int KC = TK / 8;
    constexpr int B_BYTES = BN * TK * 2;
    constexpr int U = 4;                           // loads in flight per loader thread
    __shared__ __align__(8) uint64_t raw_full[MAXNG];            // TMA path: the raw halo of a chunk has landed (byte count)
    __shared__ __align__(8) uint64_t a_full[MAXNG], a_empty[MAXNG], w_full[NB], w_empty[NB], acc_full[MAXACC], acc_empty[MAXACC];
    __shared__ uint32_t tmem_slot;
    __shared__ int s_pix[PIXTAB];                  // one pixel table per loader group (ng x NPX <= PIXTAB)
    __shared__ int s_lut[MAXNPX];                  // halo pixel -> (plane << 20 | row << 10 | column), tile independent
    __shared__ __align__(16) float s_bias[256];                  // bias of the 1x1 heads (Cout <= 256), zero padded to whole n-tiles
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool PROF = (A.dbg & 8) != 0;
    long long prof_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    const long long prof_start = clock64();
    const int a_bytes = A.a_bytes;
    unsigned char* sEpi = smem;                                  // [EPI_SMEM] epilogue scratch (never aliased)
    unsigned char* sA = smem + EPI_SMEM;                         // ng halo buffers
    unsigned char* sB = sA + A.ng * a_bytes;                     // NB weight slots

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < MAXNG; ++i) { mbar_init(&a_full[i], (uint32_t)(LOADERS / A.ng)); mbar_init(&a_empty[i], 1); mbar_init(&raw_full[i], 1); }
        if (A.use_tma) { tma_prefetch_desc(&tm0); if (A.nsrc > 1) tma_prefetch_desc(&tm1); }
#pragma unroll
        for (int i = 0; i < NB; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
#pragma unroll
        for (int i = 0; i < MAXACC; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], EPI); }
        fence_mbar_init();
    }
    if (warp == 12) tmem_alloc(&tmem_slot, (uint32_t)A.tmem_cols);
    {
        const int ppl = A.PH * A.PW;
        for (int h = tid; h < A.NPX; h += CTA) {
            const int p = h / ppl; const int r = h - p * ppl; const int hy = r / A.PW; const int hx = r - hy * A.PW;
            s_lut[h] = (p << 20) | (hy << 10) | hx;
        }
    }
    for (int i = tid; i < 256; i += CTA) s_bias[i] = (A.bias && i < A.Cout) ? A.bias[i] : 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;

    if (warp >= 4 && warp < 12) {
        // ------------------------------------------------------------------ halo loaders
        // The 256 loader threads form ng groups; group gi owns halo buffer gi and fills it with the chunks gi, gi + ng, ...
        // of this CTA's (tile, K chunk) sequence, so ng gathers (global-load round trips) are in flight at once.
        const int GT = LOADERS / A.ng;
        const int gi = (tid - EPI) / GT, lt = (tid - EPI) - gi * GT;
        const int kc = lt % KC, h0 = lt / KC;
        const int HSTEP = GT / KC;                               // halo pixels covered by one pass of a group
        int* my_pix = s_pix + gi * A.NPX;
        const int my_tiles = (A.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int total_chunks = my_tiles * A.nkt;
        if (A.use_tma) {
            // ---- tiled TMA: lane 0 of the group issues one 5-D box per parity plane (8 channels x PW pixels x PH rows x KC cores,
            // element strides = the convolution stride, zero fill outside the image); the group's threads then normalise the
            // landed data in place (nothing to do for a source without a producer BatchNorm) and hand the buffer to the MMA warp
            const int ppl = A.PH * A.PW;
            int cur_seq = -1;
            TileCoord tc_ = decode_tile(A, blockIdx.x);
            bool interior = true;
            for (int cc = gi; cc < total_chunks; cc += A.ng) {
                const int tseq = fdiv(cc, A.nkt, A.inv_nkt);
                const int c = cc - tseq * A.nkt;
                if (tseq != cur_seq) {
                    cur_seq = tseq;
                    tc_ = decode_tile(A, (int)blockIdx.x + tseq * (int)gridDim.x);
                    interior = true;
                    for (int p = 0; p < A.nplane; ++p) {
                        const int y0 = (tc_.a0 + A.plane[p].oy) * A.istr + A.plane[p].qy, x0 = (tc_.b0 + A.plane[p].ox) * A.istr + A.plane[p].qx;
                        interior = interior && y0 >= 0 && x0 >= 0 && y0 + (A.PH - 1) * A.istr < A.Hin && x0 + (A.PW - 1) * A.istr < A.Win;
                    }
                }
                const int g = tc_.g;
                const int b = gi;
                const uint32_t use = (uint32_t)(cc / A.ng);
                const int si = c < A.nkt0 ? 0 : 1;
                const rp_conv_src& S = A.src[si];
                const int ch = (c - (si ? A.nkt0 : 0)) * TK;
                const bool act = S.act != 0;
                rp_h162 sc2[4], sh2[4];
                if (act) {                                   // issued before the waits below: their latency is hidden
                    const int kc_ = lt / (GT / KC);
                    const float4* ps = reinterpret_cast<const float4*>(S.scale + (size_t)g * S.sstride + S.s_off + ch + kc_ * 8);
                    const float4* ph = reinterpret_cast<const float4*>(S.shift + (size_t)g * S.sstride + S.s_off + ch + kc_ * 8);
                    const float4 s0 = __ldg(ps), s1 = __ldg(ps + 1), q0 = __ldg(ph), q1 = __ldg(ph + 1);
                    sc2[0] = rp_f2_to_h2(s0.x, s0.y); sc2[1] = rp_f2_to_h2(s0.z, s0.w); sc2[2] = rp_f2_to_h2(s1.x, s1.y); sc2[3] = rp_f2_to_h2(s1.z, s1.w);
                    sh2[0] = rp_f2_to_h2(q0.x, q0.y); sh2[1] = rp_f2_to_h2(q0.z, q0.w); sh2[2] = rp_f2_to_h2(q1.x, q1.y); sh2[3] = rp_f2_to_h2(q1.z, q1.w);
                }
                HP_T0(w0)
                mbar_wait_backoff(&a_empty[b], (use & 1u) ^ 1u);
                HP_ADD(1, w0)
                unsigned char* buf = sA + b * a_bytes;
                HP_T0(i0)
                if (lt == 0) {
                    mbar_expect_tx(&raw_full[b], (uint32_t)(A.nplane * KC * ppl * 16));
                    for (int p = 0; p < A.nplane; ++p) {
                        const int y0 = (tc_.a0 + A.plane[p].oy) * A.istr + A.plane[p].qy, x0 = (tc_.b0 + A.plane[p].ox) * A.istr + A.plane[p].qx;
                        tma_load_5d(buf + p * A.plane_bytes, si ? (const void*)&tm1 : (const void*)&tm0, 0, x0, y0, ch >> 3, tc_.img, &raw_full[b]);
                    }
                }
                HP_ADD(13, i0)
                HP_T0(c0)
                mbar_wait(&raw_full[b], use & 1u);
                HP_ADD(2, c0)
                HP_T0(x0t)
                if (act && !(A.dbg & 2)) {
                    // thread -> one K core kc (scale / shift of its 8 channels stay in registers) and the plane pixels
                    // h = j, j + TPK, ... (TPK = GT / KC threads per core).  The TMA box is dense, so the K cores cannot be skewed
                    // against the banks the way the gather layout is: with only four threads per core the two cores that share
                    // a quarter-warp take alternating halves of every 8-pixel run instead, which keeps LDS.128 / STS.128 conflict-free
                    const int TPK = GT / KC, kc = lt / TPK, j = lt - kc * TPK;
                    const rp_h162 sl2 = rp_f2_to_h2(S.slope, S.slope);
                    for (int p = 0; p < A.nplane; ++p) {
                        unsigned char* base = buf + p * A.plane_bytes + (size_t)kc * ppl * 16;
                        const int py0 = (tc_.a0 + A.plane[p].oy) * A.istr + A.plane[p].qy, px0 = (tc_.b0 + A.plane[p].ox) * A.istr + A.plane[p].qx;
                        const int nit = TPK >= 8 ? (ppl - j + TPK - 1) / TPK : 2 * ((ppl + 7) / 8);
                        for (int ib = 0; ib < nit; ib += 4) {
                            uint4 x[4]; bool live[4]; int hs[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int i = ib + u;
                                const int h = TPK >= 8 ? j + TPK * i : j + 4 * ((kc + i) & 1) + 8 * (i >> 1);
                                live[u] = i < nit && h < ppl;
                                hs[u] = live[u] ? h : 0;
                                if (live[u] && !interior) {          // zero padding stays zero (the BatchNorm shift must not leak in)
                                    const int l = s_lut[h];
                                    const int iy = py0 + ((l >> 10) & 1023) * A.istr, ix = px0 + (l & 1023) * A.istr;
                                    live[u] = iy >= 0 && iy < A.Hin && ix >= 0 && ix < A.Win;
                                }
                                x[u] = *reinterpret_cast<const uint4*>(base + (size_t)hs[u] * 16);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                rp_h162* hp = reinterpret_cast<rp_h162*>(&x[u]);
#pragma unroll
                                for (int q = 0; q < 4; ++q) { const rp_h162 z = __hfma2(hp[q], sc2[q], sh2[q]); hp[q] = __hmax2(z, __hmul2(z, sl2)); }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                if (live[u]) *reinterpret_cast<uint4*>(base + (size_t)hs[u] * 16) = x[u];
                            }
                        }
                    }
                    fence_async_smem();            // generic-proxy writes -> visible to the tensor core (async proxy)
                }
                HP_ADD(3, x0t)
                mbar_arrive(&a_full[b]);
            }
        } else {
        int cur_seq = -1;
        TileCoord tc_ = decode_tile(A, blockIdx.x);
        for (int cc = gi; cc < total_chunks; cc += A.ng) {
            const int tseq = fdiv(cc, A.nkt, A.inv_nkt);
            const int c = cc - tseq * A.nkt;
            if (tseq != cur_seq) {
                cur_seq = tseq;
                tc_ = decode_tile(A, (int)blockIdx.x + tseq * (int)gridDim.x);
                HP_T0(tb0)
                named_bar_sync(2 + gi, GT);                      // the group is done with its previous table
                for (int h = lt; h < A.NPX; h += GT) {           // global pixel index of every halo pixel (-1 = zero padding)
                    const int l = s_lut[h];
                    const int p = l >> 20, hy = (l >> 10) & 1023, hx = l & 1023;
                    const int iy = (tc_.a0 + A.plane[p].oy + hy) * A.istr + A.plane[p].qy;
                    const int ix = (tc_.b0 + A.plane[p].ox + hx) * A.istr + A.plane[p].qx;
                    my_pix[h] = (iy >= 0 && iy < A.Hin && ix >= 0 && ix < A.Win) ? (tc_.img * A.Hin + iy) * A.Win + ix : -1;
                }
                named_bar_sync(2 + gi, GT);
                HP_ADD(12, tb0)
            }
            {
                const int g = tc_.g;
                const int b = gi;
                const uint32_t use = (uint32_t)(cc / A.ng);      // how often this buffer has been filled before
                const int si = c < A.nkt0 ? 0 : 1;
                const rp_conv_src& S = A.src[si];
                const int ch = (c - (si ? A.nkt0 : 0)) * TK + kc * 8;
                float sc[8], sh[8];
                const bool act = S.act != 0;
                const float slope = S.slope;
                if (act) {
                    const float4* ps = reinterpret_cast<const float4*>(S.scale + (size_t)g * S.sstride + S.s_off + ch);
                    const float4* ph = reinterpret_cast<const float4*>(S.shift + (size_t)g * S.sstride + S.s_off + ch);
                    const float4 s0 = __ldg(ps), s1 = __ldg(ps + 1), q0 = __ldg(ph), q1 = __ldg(ph + 1);
                    sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
                    sh[0] = q0.x; sh[1] = q0.y; sh[2] = q0.z; sh[3] = q0.w; sh[4] = q1.x; sh[5] = q1.y; sh[6] = q1.z; sh[7] = q1.w;
                } else {
#pragma unroll
                    for (int q = 0; q < 8; ++q) { sc[q] = 1.f; sh[q] = 0.f; }
                }
                const bool in_bf16 = S.dtype == 1;
                const size_t esz = in_bf16 ? 2 : 4;
                const unsigned char* base = reinterpret_cast<const unsigned char*>(S.ptr) + (size_t)(S.ch_off + ch) * esz;
                const size_t pstride = (size_t)S.pitch * esz;
                HP_T0(w0)
                mbar_wait_backoff(&a_empty[b], (use & 1u) ^ 1u); // the MMAs that read this buffer are done
                HP_ADD(1, w0)
                unsigned char* dst = sA + b * a_bytes + kc * A.a_lbo;
                auto finish = [&](float (&v)[8], int h) {       // BatchNorm + LeakyReLU, bf16, one 16-byte core row
                    if (act) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) { const float z = fmaf(v[q], sc[q], sh[q]); v[q] = z > 0.f ? z : slope * z; }
                    }
                    if (SPLIT) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) v[q] *= A.in_scale;       // 2^4: the low halves of |x| > 0.008 stay normal numbers
                        if (A.fused3 || A.split_lo) {
                            float l[8];
#pragma unroll
                            for (int q = 0; q < 8; q += 2) {
                                const float2 hi = rp_h2_to_f2(rp_f2_to_h2(v[q], v[q + 1]));
                                l[q] = v[q] - hi.x; l[q + 1] = v[q + 1] - hi.y;
                            }
                            if (A.fused3) {                     // the low halves go to the second halo of the buffer
                                rp_h162 l0 = rp_f2_to_h2(l[0], l[1]), l1 = rp_f2_to_h2(l[2], l[3]), l2 = rp_f2_to_h2(l[4], l[5]), l3 = rp_f2_to_h2(l[6], l[7]);
                                uint4 ol;
                                ol.x = *reinterpret_cast<uint32_t*>(&l0); ol.y = *reinterpret_cast<uint32_t*>(&l1);
                                ol.z = *reinterpret_cast<uint32_t*>(&l2); ol.w = *reinterpret_cast<uint32_t*>(&l3);
                                *reinterpret_cast<uint4*>(dst + A.a_half + (size_t)h * 16) = ol;
                            } else {                            // this launch multiplies the low halves
#pragma unroll
                                for (int q = 0; q < 8; ++q) v[q] = l[q];
                            }
                        }
                    }
                    rp_h162 p0 = rp_f2_to_h2(v[0], v[1]), p1 = rp_f2_to_h2(v[2], v[3]);
                    rp_h162 p2 = rp_f2_to_h2(v[4], v[5]), p3 = rp_f2_to_h2(v[6], v[7]);
                    uint4 o;
                    o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
                    o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
                    *reinterpret_cast<uint4*>(dst + (size_t)h * 16) = o;
                };
                long long x0 = 0;
                if (A.dbg & 2) {
                } else if (in_bf16) {
                    // 16-bit sources: the raw halo goes global -> shared with cp.async (LDGSTS: no registers held per load),
                    // so a loader group has its WHOLE K chunk in flight instead of 8 loads per thread -- these layers are
                    // bound by memory-level parallelism.  Padding pixels are zero-filled by the copy itself.  Each thread then
                    // normalises, in place, exactly the units it copied (no cross-thread hand-off); a source without a
                    // producer BatchNorm (the split stem input) needs no second pass at all.
                    HP_T0(i0)
                    for (int h = h0; h < A.NPX; h += HSTEP) {
                        const int pix = my_pix[h];
                        cp_async16(dst + (size_t)h * 16, base + (size_t)(pix < 0 ? 0 : pix) * pstride, pix < 0 ? 0u : 16u);
                    }
                    HP_ADD(13, i0)
                    HP_T0(c0)
                    cp_async_wait_all();
                    HP_ADD(2, c0)
                    if (PROF) x0 = clock64();
                    if (act && A.h2math && !(SPLIT && A.split_lo)) {
                        // packed half arithmetic: z = x * scale + shift is ONE rounding of the exact value for half inputs (the
                        // float path rounds to half after the float FMA as well); what differs is scale / shift rounded to half.
                        // LeakyReLU(z) = max(z, slope z) for 0 <= slope <= 1.  12 instructions per 8 channels instead of ~36.
                        rp_h162 sc2[4], sh2[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) { sc2[q] = rp_f2_to_h2(sc[2 * q], sc[2 * q + 1]); sh2[q] = rp_f2_to_h2(sh[2 * q], sh[2 * q + 1]); }
                        const rp_h162 sl2 = rp_f2_to_h2(slope, slope);
                        // four units per step, straight-line (loads first, then the math, then the stores): padding units are
                        // recomputed as zero through a select instead of a branch so that the four chains interleave
                        for (int hb = h0; hb < A.NPX; hb += 4 * HSTEP) {
                            uint4 x[4]; bool live[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int h = hb + u * HSTEP;
                                live[u] = h < A.NPX && my_pix[h < A.NPX ? h : 0] >= 0;
                                x[u] = *reinterpret_cast<const uint4*>(dst + (size_t)(h < A.NPX ? h : h0) * 16);
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                rp_h162* hp = reinterpret_cast<rp_h162*>(&x[u]);
#pragma unroll
                                for (int q = 0; q < 4; ++q) { const rp_h162 z = __hfma2(hp[q], sc2[q], sh2[q]); hp[q] = __hmax2(z, __hmul2(z, sl2)); }
                            }
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int h = hb + u * HSTEP;
                                if (live[u]) *reinterpret_cast<uint4*>(dst + (size_t)h * 16) = x[u];
                            }
                        }
                    } else if (act || (SPLIT && A.split_lo)) {
#pragma unroll 2
                        for (int h = h0; h < A.NPX; h += HSTEP) {
                            if (my_pix[h] < 0) continue;              // zero padding stays zero (the BatchNorm shift must not leak in)
                            const uint4 x = *reinterpret_cast<const uint4*>(dst + (size_t)h * 16);
                            float v[8];
                            const rp_h162* hp = reinterpret_cast<const rp_h162*>(&x);
#pragma unroll
                            for (int q = 0; q < 4; ++q) { const float2 f = rp_h2_to_f2(hp[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
                            finish(v, h);
                        }
                    }
                } else {
                    for (int hb = h0; hb < A.NPX; hb += HSTEP * U) {
                        uint4 x[U][2];
                        int pix[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int h = hb + u * HSTEP;
                            pix[u] = h < A.NPX ? my_pix[h] : -2;
                            if (pix[u] >= 0) {
                                const uint4* p = reinterpret_cast<const uint4*>(base + (size_t)pix[u] * pstride);
                                x[u][0] = __ldg(p);
                                x[u][1] = __ldg(p + 1);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (pix[u] == -2) continue;
                            const int h = hb + u * HSTEP;
                            if (pix[u] < 0) {
                                *reinterpret_cast<uint4*>(dst + (size_t)h * 16) = make_uint4(0u, 0u, 0u, 0u);
                                if (SPLIT && A.fused3) *reinterpret_cast<uint4*>(dst + A.a_half + (size_t)h * 16) = make_uint4(0u, 0u, 0u, 0u);
                                continue;
                            }
                            float v[8] = {__uint_as_float(x[u][0].x), __uint_as_float(x[u][0].y), __uint_as_float(x[u][0].z), __uint_as_float(x[u][0].w),
                                          __uint_as_float(x[u][1].x), __uint_as_float(x[u][1].y), __uint_as_float(x[u][1].z), __uint_as_float(x[u][1].w)};
                            finish(v, h);
                        }
                    }
                }
                if (in_bf16 && !(A.dbg & 2)) { HP_ADD(3, x0) }
                HP_T0(f0)
                fence_async_smem();                // generic-proxy writes -> visible to the tensor core (async proxy)
                mbar_arrive(&a_full[b]);
                HP_ADD(14, f0)
            }
        }
        }
    } else if (warp == 12) {
        // ------------------------------------------------------------------ weight stream (bulk TMA)
        // (the whole warp runs the loop so that its control values stay warp-uniform; one elected lane issues)
        const bool leader = elect_one();
        const int per_tile = A.nkt * A.ntap * ((SPLIT && (A.fused3 || A.pairw)) ? 2 : 1);    // split precision: (hi, lo) block pairs
        uint32_t wi = 0;
        if (A.w_resident) {
            // small layers (stem, heads, 32/64-channel 4x4 layers with one n-tile): the whole weight set sits in the ring
            if (leader) {
                for (int i = 0; i < per_tile; ++i) {
                    mbar_expect_tx(&w_full[i], B_BYTES);
                    bulk_g2s(sB + i * B_BYTES, Wp + (size_t)i * B_BYTES, B_BYTES, &w_full[i]);
                }
            }
            __syncwarp();
        } else {
            for (int tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x) {
                const int tile_n = tile - fdiv(tile, A.ntn, A.inv_ntn) * A.ntn;
                const unsigned char* wsrc = Wp + (size_t)tile_n * per_tile * B_BYTES;
                for (int i = 0; i < per_tile; ++i, ++wi) {
                    const int slot = wi % NB;
                    mbar_wait_backoff(&w_empty[slot], (uint32_t)(((wi / NB) & 1) ^ 1));
                    if (leader) {
                        mbar_expect_tx(&w_full[slot], B_BYTES);
                        bulk_g2s(sB + slot * B_BYTES, wsrc + (size_t)i * B_BYTES, B_BYTES, &w_full[slot]);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 13) {
        // ------------------------------------------------------------------ MMA issue
        // Converged warp; descriptors are built from warp-uniform 32-bit words: low = start address >> 4 | LBO >> 4 << 16
        // (a tap / K step only adds to the start-address field), high = SBO >> 4 | version 1 << 14.
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_bf16(TM, BN);
        const uint32_t a_hi = (((uint32_t)A.PW * 16u) >> 4) | (1u << 14);
        const uint32_t b_hi = (128u >> 4) | (1u << 14);
        const uint32_t a_lo_k = (((uint32_t)A.a_lbo >> 4) & 0x3fffu) << 16;
        const uint32_t b_lo_k = ((uint32_t)((BN / 8) * 128) >> 4) << 16;
        const uint32_t a_step = (uint32_t)(2 * A.a_lbo) >> 4;              // one K=16 step = 2 K cores
        constexpr uint32_t b_step = (uint32_t)(2 * (BN / 8) * 128) >> 4;
        const uint32_t sA16 = smem_u32(sA) >> 4, sB16 = smem_u32(sB) >> 4;
        // per-tap constants in registers (the tap loops below are fully unrolled; taps beyond ntap are skipped)
        uint32_t tap_a[MAXT], tap_col[MAXT];
        uint32_t first_mask = 0;
#pragma unroll
        for (int t = 0; t < MAXT; ++t) {
            tap_a[t] = t < A.ntap ? ((uint32_t)A.tap[t].a_off >> 4) : 0u;
            tap_col[t] = t < A.ntap ? (uint32_t)(A.tap[t].cls * BN) : 0u;
            if (t < A.ntap && A.tap[t].first) first_mask |= 1u << t;
        }
        uint32_t wi = 0, cc = 0, tl = 0;
        bool w_ready = false;
        for (int tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, ++tl) {
            const uint32_t ab = tl % (uint32_t)A.nacc;
            HP_T0(m0)
            mbar_wait(&acc_empty[ab], ((tl / (uint32_t)A.nacc) & 1u) ^ 1u);   // the epilogue has drained this accumulator set
            HP_ADD(5, m0)
            tc_fence_after();
            const uint32_t acc0 = tmem_d + ab * (uint32_t)A.acc_cols;
            const bool F3 = SPLIT && (A.fused3 || A.pairw);          // weight blocks come as (hi, lo) pairs
            const bool TWOHALO = SPLIT && A.fused3;
            if (A.w_resident && !w_ready) {                               // resident weights: wait for them once
                for (int i = 0; i < A.nkt * A.ntap * (F3 ? 2 : 1); ++i) mbar_wait(&w_full[i], 0u);
                w_ready = true;
            }
            for (int c = 0; c < A.nkt; ++c, ++cc) {
                const uint32_t b = cc % (uint32_t)A.ng;
                HP_T0(m1)
                mbar_wait(&a_full[b], (cc / (uint32_t)A.ng) & 1u);
                HP_ADD(6, m1)
                tc_fence_after();
                const uint32_t a_lo_buf = a_lo_k | (sA16 + b * ((uint32_t)a_bytes >> 4));
                const uint32_t fresh_mask = c == 0 ? first_mask : 0u;
                if (F3) {
                    // split precision: per tap the weight blocks (hi, lo); A_hi B_hi (+ A_lo B_hi with two halos) + A_hi B_lo into ONE accumulator
                    const uint32_t a_lo_buf2 = a_lo_buf + ((uint32_t)A.a_half >> 4);
#pragma unroll
                    for (int t = 0; t < MAXT; ++t) {
                        if (t < A.ntap) {
                            const uint32_t a_h = a_lo_buf + tap_a[t], a_l = a_lo_buf2 + tap_a[t];
                            uint32_t slot_h, slot_l;
                            if (A.w_resident) { slot_h = (uint32_t)(c * A.ntap + t) * 2u; slot_l = slot_h + 1u; }
                            else {
                                slot_h = wi % NB; slot_l = (wi + 1) % NB;
                                HP_T0(m2)
                                mbar_wait(&w_full[slot_h], (uint32_t)((wi / NB) & 1));
                                mbar_wait(&w_full[slot_l], (uint32_t)(((wi + 1) / NB) & 1));
                                HP_ADD(7, m2)
                            }
                            const uint32_t b_h = b_lo_k | (sB16 + slot_h * (uint32_t)(B_BYTES >> 4));
                            const uint32_t b_l = b_lo_k | (sB16 + slot_l * (uint32_t)(B_BYTES >> 4));
                            if (leader) {
#pragma unroll
                                for (int j = 0; j < TK / 16; ++j) {
                                    umma_bf16(acc0 + tap_col[t], pack_desc(a_h + j * a_step, a_hi), pack_desc(b_h + j * b_step, b_hi), idesc,
                                              (j > 0 || !((fresh_mask >> t) & 1u)) ? 1u : 0u);
                                    if (TWOHALO) umma_bf16(acc0 + tap_col[t], pack_desc(a_l + j * a_step, a_hi), pack_desc(b_h + j * b_step, b_hi), idesc, 1u);
                                    umma_bf16(acc0 + tap_col[t], pack_desc(a_h + j * a_step, a_hi), pack_desc(b_l + j * b_step, b_hi), idesc, 1u);
                                }
                                if (!A.w_resident) { umma_commit(&w_empty[slot_h]); umma_commit(&w_empty[slot_l]); }
                            }
                            if (!A.w_resident) wi += 2;
                        }
                    }
                } else if (A.w_resident) {
                    const uint32_t b_lo0 = b_lo_k | (sB16 + (uint32_t)(c * A.ntap) * (uint32_t)(B_BYTES >> 4));
#pragma unroll
                    for (int t = 0; t < MAXT; ++t) {
                        if (t < A.ntap) {
                            const uint32_t a_lo = a_lo_buf + tap_a[t];
                            const uint32_t b_lo = b_lo0 + (uint32_t)t * (uint32_t)(B_BYTES >> 4);
                            if (leader && !(A.dbg & 4)) {
#pragma unroll
                                for (int j = 0; j < TK / 16; ++j)
                                    umma_bf16(acc0 + tap_col[t], pack_desc(a_lo + j * a_step, a_hi), pack_desc(b_lo + j * b_step, b_hi), idesc,
                                              (j > 0 || !((fresh_mask >> t) & 1u)) ? 1u : 0u);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int t = 0; t < MAXT; ++t) {
                        if (t < A.ntap) {
                            const uint32_t slot = wi % NB;
                            HP_T0(m2)
                            mbar_wait(&w_full[slot], (uint32_t)((wi / NB) & 1));
                            HP_ADD(7, m2)
                            const uint32_t a_lo = a_lo_buf + tap_a[t];
                            const uint32_t b_lo = b_lo_k | (sB16 + slot * (uint32_t)(B_BYTES >> 4));
                            if (leader) {
                                if (!(A.dbg & 4)) {
#pragma unroll
                                    for (int j = 0; j < TK / 16; ++j)
                                        umma_bf16(acc0 + tap_col[t], pack_desc(a_lo + j * a_step, a_hi), pack_desc(b_lo + j * b_step, b_hi), idesc,
                                                  (j > 0 || !((fresh_mask >> t) & 1u)) ? 1u : 0u);
                                }
                                umma_commit(&w_empty[slot]);   // frees the weight slot once these MMAs have read it
                            }
                            ++wi;
                        }
                    }
                }
                if (leader) umma_commit(&a_empty[b]);  // frees the halo buffer
                __syncwarp();
            }
            if (leader) umma_commit(&acc_full[ab]);    // accumulator set complete
            __syncwarp();
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 0-3 = TMEM lane quarters)
        float* Tt = reinterpret_cast<float*>(sEpi) + warp * (32 * 33);         // per-warp [32 rows][33]
        float* red = reinterpret_cast<float*>(sEpi) + 4 * 32 * 33;             // [unit][quarter][32] x {sum, sq}
        const int q = warp;
        const int m = q * 32 + lane;
        const int nunit = A.nclass * (BN / 32);
        uint32_t tl = 0;
        for (int tile = blockIdx.x; tile < A.total_tiles; tile += gridDim.x, ++tl) {
            const TileCoord tc_ = decode_tile(A, tile);
            const uint32_t ab = tl % (uint32_t)A.nacc;
            const int a = tc_.a0 + (m >> 3), bcol = tc_.b0 + (m & 7);
            HP_T0(e0)
            mbar_wait(&acc_full[ab], (tl / (uint32_t)A.nacc) & 1u);
            HP_ADD(9, e0)
            tc_fence_after();
            for (int u = 0; u < nunit; ++u) {
                const int cls = u / (BN / 32), c0 = (u - cls * (BN / 32)) * 32;
                float v[32];
                HP_T0(e1)
                tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + ab * (uint32_t)A.acc_cols + (uint32_t)(cls * BN + c0), v);
                HP_ADD(10, e1)
                const bool valid = a < A.cls_Ha[cls] && bcol < A.cls_Wb[cls];
                const int co0 = tc_.tile_n * BN + c0;
                if (SPLIT) {                                // activations and weights went in scaled by 2^4 / 2^8
                    const float os = A.out_scale;
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= os;
                }
                if (SPLIT && A.accum && valid) {            // a later launch of the split-precision mode: add onto what is stored
                    const int oy_ = a * A.ostr + A.cls_py[cls], ox_ = bcol * A.ostr + A.cls_px[cls];
                    const float* pp = reinterpret_cast<const float*>(A.out) + (((size_t)tc_.img * A.Hout + oy_) * A.Wout + ox_) * A.out_pitch + A.out_ch_off + co0;
                    if (co0 + 31 < A.Cout && ((((size_t)pp) & 15) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 p4 = *reinterpret_cast<const float4*>(pp + j);
                            v[j] += p4.x; v[j + 1] += p4.y; v[j + 2] += p4.z; v[j + 3] += p4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (co0 + j < A.Cout) v[j] += pp[j];
                    }
                }
                if (A.bias || A.tanh_out) {                 // (bias/tanh layers have Cout <= 256, checked on the host)
                    const float4* bp = reinterpret_cast<const float4*>(s_bias + co0);
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = bp[j >> 2];
                        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                    }
                    if (A.tanh_out) {
                        if (SPLIT) {                            // split-precision modes: float32-class output, so the exact function
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = tanhf(v[j]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = fast_tanh(v[j]);
                        }
                    }
                }
                rp_h162 pk[16];            // 16-bit output: packed once, stored as packed
                if (A.out_bf16) {          // statistics describe what the consumer will read: the rounded values
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        pk[j] = rp_f2_to_h2(v[2 * j], v[2 * j + 1]);
                        const float2 f = rp_h2_to_f2(pk[j]);
                        v[2 * j] = f.x; v[2 * j + 1] = f.y;
                    }
                }
                if (valid && !(A.dbg & 1)) {
                    const int oy = a * A.ostr + A.cls_py[cls], ox = bcol * A.ostr + A.cls_px[cls];
                    const size_t opix = ((size_t)tc_.img * A.Hout + oy) * A.Wout + ox;
                    if (A.out_bf16) {
                        rp_h16* op = reinterpret_cast<rp_h16*>(A.out) + opix * A.out_pitch + A.out_ch_off + co0;
                        if (co0 + 31 < A.Cout) {
#pragma unroll
                            for (int j = 0; j < 32; j += 8) {
                                uint4 o;
                                o.x = *reinterpret_cast<uint32_t*>(&pk[j / 2]); o.y = *reinterpret_cast<uint32_t*>(&pk[j / 2 + 1]);
                                o.z = *reinterpret_cast<uint32_t*>(&pk[j / 2 + 2]); o.w = *reinterpret_cast<uint32_t*>(&pk[j / 2 + 3]);
                                *reinterpret_cast<uint4*>(op + j) = o;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (co0 + j < A.Cout) op[j] = rp_f_to_h(v[j]);
                        }
                    } else {
                        float* op = reinterpret_cast<float*>(A.out) + opix * A.out_pitch + A.out_ch_off + co0;
                        if (co0 + 31 < A.Cout && ((((size_t)op) & 15) == 0)) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) if (co0 + j < A.Cout) op[j] = v[j];
                        }
                    }
                }
                if (A.psum && !(A.dbg & 1)) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) Tt[lane * 33 + j] = valid ? v[j] : 0.f;
                    __syncwarp();
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int r = 0; r < 32; ++r) { const float x = Tt[r * 33 + lane]; s1 += x; s2 += x * x; }
                    red[(u * 4 + q) * 64 + lane] = s1;
                    red[(u * 4 + q) * 64 + 32 + lane] = s2;
                    __syncwarp();
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[ab]);               // this thread's tcgen05.ld of the set have completed (wait::ld)
            HP_T0(e2)
            if (A.psum) {
                named_bar_sync(14, EPI);
                for (int i = tid; i < nunit * 32; i += EPI) {
                    const int u = i >> 5, col = i & 31;
                    const int cls = u / (BN / 32), c0 = (u - cls * (BN / 32)) * 32;
                    const int co = tc_.tile_n * BN + c0 + col;
                    if (co < A.Cout) {
                        const float* r = red + (u * 4) * 64;
                        const float s1 = ((r[col] + r[64 + col]) + r[128 + col]) + r[192 + col];
                        const float s2 = ((r[32 + col] + r[96 + col]) + r[160 + col]) + r[224 + col];
                        const size_t prow = (size_t)(tc_.g * A.nclass + cls) * A.tiles_m + tc_.tile_m;
                        A.psum[prow * A.Cout + co] = s1;
                        A.psq[prow * A.Cout + co] = s2;
                    }
                }
                named_bar_sync(14, EPI);               // `red` is free for the next tile
            }
            HP_ADD(11, e2)
        }
    }
    if (PROF && lane == 0) {
        const long long tot = clock64() - prof_start;
        if (warp == 4) { atomicAdd(&g_halo_prof[0], (unsigned long long)tot); for (int i = 1; i <= 3; ++i) atomicAdd(&g_halo_prof[i], (unsigned long long)prof_acc[i]); atomicAdd(&g_halo_prof[12], (unsigned long long)prof_acc[12]); atomicAdd(&g_halo_prof[13], (unsigned long long)prof_acc[13]); atomicAdd(&g_halo_prof[14], (unsigned long long)prof_acc[14]); }
        if (warp == 13) { atomicAdd(&g_halo_prof[4], (unsigned long long)tot); for (int i = 5; i <= 7; ++i) atomicAdd(&g_halo_prof[i], (unsigned long long)prof_acc[i]); }
        if (warp == 0) { atomicAdd(&g_halo_prof[8], (unsigned long long)tot); for (int i = 9; i <= 11; ++i) atomicAdd(&g_halo_prof[i], (unsigned long long)prof_acc[i]); }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 12) tmem_dealloc(tmem_d, (uint32_t)A.tmem_cols);
}

// -------------------------------------------------------------------------------------------------------------------
static inline int floordiv(int a, int b) { int q = a / b; if ((a % b != 0) && ((a < 0) != (b < 0))) --q; return q; }

// Tile plan of a layer; `pitch16` pads the halo row pitch to 16 pixels (SBO multiple of 256 B; debugging aid).
static bool build_halo_args(const rp_conv_desc* d, HaloArgs* H, int bn, int tk, int* tap_widx, bool pitch16) {
    scnet::ConvArgs A;
    if (!scnet::build_args(d, &A, TM)) return false;
    H->nsrc = A.nsrc;
    for (int i = 0; i < A.nsrc; ++i) {
        H->src[i] = A.src[i];
        const int al = A.src[i].dtype == 1 ? 8 : 4;
        if (A.src[i].C % tk || A.src[i].pitch % al || A.src[i].ch_off % al) return false;
        if (A.src[i].act && ((A.src[i].sstride % 4) || (A.src[i].s_off % 4))) return false;
    }
    H->G = A.G; H->gsz = A.gsz; H->Hin = A.Hin; H->Win = A.Win; H->Hout = A.Hout; H->Wout = A.Wout; H->Cout = A.Cout;
    H->istr = A.istr; H->ostr = A.ostr; H->nclass = A.nclass;
    int Hb = 0, Wb = 0, ntap = 0;
    for (int c = 0; c < A.nclass; ++c) {
        H->cls_py[c] = A.cls[c].py; H->cls_px[c] = A.cls[c].px; H->cls_Ha[c] = A.cls[c].Ha; H->cls_Wb[c] = A.cls[c].Wb;
        if (A.cls[c].Ha > Hb) Hb = A.cls[c].Ha;
        if (A.cls[c].Wb > Wb) Wb = A.cls[c].Wb;
        if (A.cls[c].ntap < 1) return false;
        ntap += A.cls[c].ntap;
    }
    if (ntap > MAXT || Hb < 1 || Wb < 1) return false;
    H->nty = (Hb + TH - 1) / TH; H->ntx = (Wb + TW - 1) / TW; H->tiles_m = A.gsz * H->nty * H->ntx;
    H->ntn = (A.Cout + bn - 1) / bn;
    H->total_tiles = A.G * H->tiles_m * H->ntn;
    if (H->total_tiles >= (1 << 24)) return false;
    H->inv_ntn = 1.f / (float)H->ntn; H->inv_tiles_m = 1.f / (float)H->tiles_m;
    H->inv_tiles_img = 1.f / (float)(H->nty * H->ntx); H->inv_ntx = 1.f / (float)H->ntx;
    const int s = A.istr;
    H->nplane = s * s;
    if (H->nplane > 4) return false;
    int emin_y[4], emax_y[4], emin_x[4], emax_x[4]; bool used[4] = {false, false, false, false};
    for (int c = 0; c < A.nclass; ++c)
        for (int t = 0; t < A.cls[c].ntap; ++t) {
            const int dy = A.cls[c].taps[t].dy, dx = A.cls[c].taps[t].dx;
            const int ey = floordiv(dy, s), ex = floordiv(dx, s);
            const int p = (dy - ey * s) * s + (dx - ex * s);
            if (!used[p]) { used[p] = true; emin_y[p] = emax_y[p] = ey; emin_x[p] = emax_x[p] = ex; }
            if (ey < emin_y[p]) emin_y[p] = ey; if (ey > emax_y[p]) emax_y[p] = ey;
            if (ex < emin_x[p]) emin_x[p] = ex; if (ex > emax_x[p]) emax_x[p] = ex;
        }
    int EY = 0, EX = 0;
    for (int p = 0; p < H->nplane; ++p) {
        if (!used[p]) { emin_y[p] = emax_y[p] = emin_x[p] = emax_x[p] = 0; }
        if (emax_y[p] - emin_y[p] > EY) EY = emax_y[p] - emin_y[p];
        if (emax_x[p] - emin_x[p] > EX) EX = emax_x[p] - emin_x[p];
        H->plane[p].qy = p / s; H->plane[p].qx = p % s; H->plane[p].oy = emin_y[p]; H->plane[p].ox = emin_x[p];
    }
    H->PH = TH + EY;
    H->PW = TW + EX;
    if (pitch16) H->PW = 16;
    if (H->PW < TW + EX) return false;
    H->NPX = H->nplane * H->PH * H->PW;
    if (H->NPX > MAXNPX) return false;
    int units = H->NPX;                       // K-core stride in 16-byte units: >= NPX and == 1 (mod 8)
    while (units % 8 != 1) ++units;
    H->a_lbo = units * 16;
    int k = 0;
    for (int c = 0; c < A.nclass; ++c)
        for (int t = 0; t < A.cls[c].ntap; ++t, ++k) {
            const int dy = A.cls[c].taps[t].dy, dx = A.cls[c].taps[t].dx;
            const int ey = floordiv(dy, s), ex = floordiv(dx, s);
            const int p = (dy - ey * s) * s + (dx - ex * s);
            H->tap[k].cls = c;
            H->tap[k].first = (t == 0) ? 1 : 0;
            H->tap[k].a_off = (p * H->PH * H->PW + (ey - emin_y[p]) * H->PW + (ex - emin_x[p])) * 16;
            if (tap_widx) tap_widx[k] = A.cls[c].taps[t].widx;
        }
    H->ntap = ntap;
    H->nkt = 0;
    for (int i = 0; i < A.nsrc; ++i) H->nkt += A.src[i].C / tk;
    H->nkt0 = A.src[0].C / tk;
    H->inv_nkt = 1.f / (float)H->nkt;
    H->ng = 2;
    H->w_resident = 0;
    H->acc_cols = A.nclass * bn;
    if (2 * H->acc_cols > 512) return false;
    H->nacc = 512 / H->acc_cols;                              // 2..8 accumulator sets: the MMA warp runs that many tiles ahead of
    if (H->nacc > MAXACC) H->nacc = MAXACC;                    // the epilogue, so neither waits for the other's hand-off latency
    int cols = 32;
    while (cols < H->nacc * H->acc_cols) cols <<= 1;
    H->tmem_cols = cols;
    H->bias = d->bias; H->tanh_out = d->tanh_out;
    if (d->out_dtype == 1 && (d->bias || d->tanh_out)) return false;
    if ((d->bias || d->tanh_out) && H->ntn * bn > 256) return false;
    H->out = d->out; H->out_pitch = d->out_pitch; H->out_ch_off = d->out_ch_off; H->out_bf16 = d->out_dtype == 1 ? 1 : 0;
    if (H->out_bf16 && ((d->out_pitch % 8) || (d->out_ch_off % 8))) return false;   // float32 output: any alignment (scalar stores)
    H->psum = d->psum; H->psq = d->psq;
    return true;
}

static long long g_tma_launches = 0;    // launches whose halo arrived by tiled TMA (rp_conv_halo_tma_count)

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else cudaGetLastError();
    }
    return fn;
}

// Tensor map of one 16-bit NHWC source seen as [8 channels][x][y][K core][image]: a box of 8 x PW x PH x KC x 1 elements lands in
// shared memory as [K core][row][pixel] 16-byte units -- the canonical K-major UMMA layout of the halo -- and element strides
// (istr, istr) in x / y pick one parity plane of a stride-2 convolution.  Out-of-image coordinates are zero filled.
static bool make_src_tmap(const rp_conv_src& S, const HaloArgs& H, int KC, CUtensorMap* tm) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t pitchB = (cuuint64_t)S.pitch * 2;
    cuuint64_t gdim[5] = {8, (cuuint64_t)H.Win, (cuuint64_t)H.Hin, (cuuint64_t)(S.C / 8), (cuuint64_t)H.G * (cuuint64_t)H.gsz};
    cuuint64_t gstr[4] = {pitchB, (cuuint64_t)H.Win * pitchB, 16, (cuuint64_t)H.Hin * (cuuint64_t)H.Win * pitchB};
    cuuint32_t box[5] = {8, (cuuint32_t)(H.PW * H.istr), (cuuint32_t)(H.PH * H.istr), (cuuint32_t)KC, 1};
    cuuint32_t estr[5] = {1, (cuuint32_t)H.istr, (cuuint32_t)H.istr, 1, 1};
    void* base = const_cast<unsigned char*>(reinterpret_cast<const unsigned char*>(S.ptr)) + (size_t)S.ch_off * 2;
    if (box[1] > 256 || box[2] > 256 || (reinterpret_cast<uintptr_t>(base) & 15) || (pitchB & 15)) return false;
    return enc(tm, RP_H16_FP16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BN, int TK, int NB, bool SPLIT>
static int launch_halo_impl(const HaloArgs& H0, const void* wp, int h2math, cudaStream_t stream, bool dry_run) {
    constexpr int KC = TK / 8;
    HaloArgs H = H0;
    // 16-bit sources can take the tiled-TMA loader: halo layout [plane][K core][row][pixel] with 128-byte aligned planes.  In the
    // K-major / no-swizzle UMMA layout a TMA element is one 16-byte (pixel, K core) granule, and measured on B200 (32 scan pairs,
    // profiles/r2_scnet_halo_ncu_summary.txt) that only pays for the single-tap layers (1x1 heads: -8 %); 3x3 / transposed layers
    // are even and the stride-2 convolutions, whose parity planes need element strides of 2, lose 6-20 % against the per-thread
    // cp.async gather.  Default: TMA for single-tap layers; flags bit 7 forces it everywhere, bit 6 disables it (float32
    // sources always use the gather and its [K core][plane,row,pixel] layout).
    CUtensorMap tm[2];
    memset(tm, 0, sizeof(tm));
    H.split_lo = (h2math >> 7) & 1; H.accum = (h2math >> 8) & 1; H.fused3 = (h2math >> 9) & 1; H.pairw = (h2math >> 10) & 1;
    H.a_half = 0; H.in_scale = 1.f; H.out_scale = 1.f;
    if (SPLIT) {
        for (int i = 0; i < H.nsrc; ++i) if (H.src[i].dtype == 1) return RP_ERR_UNSUPPORTED;      // float32 sources and output only
        if (H.out_bf16 || (H.fused3 && (H.split_lo || H.accum || H.pairw))) return RP_ERR_UNSUPPORTED;
        H.in_scale = 16.f; H.out_scale = 1.f / 4096.f;
    }
    H.use_tma = (((h2math >> 5) & 1) || SPLIT) ? 0 : (((h2math >> 6) & 1) || H.ntap == 1) ? 1 : 0;
    for (int i = 0; i < H.nsrc; ++i) if (H.src[i].dtype != 1) H.use_tma = 0;
    if (H.use_tma) {
        const int ppl = H.PH * H.PW;
        for (int i = 0; i < H.nsrc && H.use_tma; ++i) if (!make_src_tmap(H.src[i], H, KC, &tm[i])) H.use_tma = 0;
        if (H.use_tma) {
            H.plane_bytes = ((KC * ppl * 16 + 127) / 128) * 128;
            for (int t = 0; t < H.ntap; ++t) {
                const int u = H.tap[t].a_off / 16, p = u / ppl, rem = u - p * ppl;
                H.tap[t].a_off = p * H.plane_bytes + rem * 16;
            }
            H.a_lbo = ppl * 16;
            H.a_bytes = H.nplane * H.plane_bytes;
            H.inv_ppl = 1.f / (float)ppl;
        }
    }
    if (!H.use_tma) { H.plane_bytes = 0; H.a_bytes = ((KC * H.a_lbo + 127) / 128) * 128; }
    if (H.fused3) { H.a_half = H.a_bytes; H.a_bytes *= 2; }
    const size_t a_bytes = (size_t)H.a_bytes;
    const size_t fixed = (size_t)EPI_SMEM + (size_t)NB * BN * TK * 2;
    auto kern = conv_halo_tc<BN, TK, NB, SPLIT>;
    static size_t limit = 0;                                    // dynamic bytes a CTA of this instantiation may take
    if (limit == 0) {
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
        limit = (size_t)227 * 1024 - fa.sharedSizeBytes - 1024;
    }
    // eight / four K chunks in flight (one per loader group) when the halo is small enough, two otherwise
    H.ng = 2;
    for (int ng = MAXNG; ng > 2; ng >>= 1)
        if (H.NPX * ng <= PIXTAB && fixed + (size_t)ng * a_bytes <= limit) { H.ng = ng; break; }
    H.w_resident = (H.ntn == 1 && H.nkt * H.ntap * ((H.fused3 || H.pairw) ? 2 : 1) <= NB) ? 1 : 0;
    if ((H.fused3 || H.pairw) && NB < 2) return RP_ERR_UNSUPPORTED;
    H.h2math = h2math & 1; H.dbg = (h2math >> 1) & 15;

    const size_t smem = fixed + H.ng * a_bytes;
    if (smem > limit) return RP_ERR_UNSUPPORTED;
    if (dry_run) return RP_OK;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
    static int n_sm = 0;
    if (n_sm == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm < 1) { cudaGetLastError(); n_sm = 148; }
    }
    const int grid = H.total_tiles < n_sm ? H.total_tiles : n_sm;       // persistent: one CTA per SM
    kern<<<grid, CTA, smem, stream>>>(H, static_cast<const unsigned char*>(wp), tm[0], tm[1]);
    if (H.use_tma) ++g_tma_launches;
    ++scnet::g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

// the split-precision launches (flags bit 8 / 9) run their own instantiation: compiled into the common one, the extra paths cost
// the 16-bit layers registers (spills in three of the eight tile shapes, stem +20 %)
template <int BN, int TK, int NB>
static int launch_halo(const HaloArgs& H0, const void* wp, int h2math, cudaStream_t stream, bool dry_run = false) {
    return ((h2math >> 7) & 15) ? launch_halo_impl<BN, TK, NB, true>(H0, wp, h2math, stream, dry_run)
                               : launch_halo_impl<BN, TK, NB, false>(H0, wp, h2math, stream, dry_run);
}

}  // namespace halo

extern "C" {

// Tile plan query: partial-statistics rows per group, number of (class, tap) weight blocks per K chunk and their order
// (tap_widx[i] = ky*k+kx of block i; the packed weights are [n-tile][K chunk][block][tk/8][bn/8][8 co][8 ci] bf16).
int rp_conv_halo_plan(const rp_conv_desc* d, int bn, int tk, int flags, int* nparts, int* ntap, int* tap_widx) {
    halo::HaloArgs H;
    if (!d || !halo::build_halo_args(d, &H, bn, tk, tap_widx, (flags & 1) != 0)) return RP_ERR_UNSUPPORTED;
    if (nparts) *nparts = H.nclass * H.tiles_m;
    if (ntap) *ntap = H.ntap;
    return RP_OK;
}

// Test hook: the tile plan as integers (tests/test_halo_plan.py emulates the kernel's addressing on the CPU with it):
// [istr, ostr, nclass, nty, ntx, tiles_m, nplane, PH, PW, NPX, a_lbo, ntap, nkt, nkt0, tmem_cols, 0] +
// 4 x plane {qy,qx,oy,ox} + 16 x tap {cls,a_off,first} + 4 x class {py,px,Ha,Wb}  (96 ints)
int rp_conv_halo_debug(const rp_conv_desc* d, int bn, int tk, int flags, int* out96) {
    halo::HaloArgs H;
    if (!d || !out96 || !halo::build_halo_args(d, &H, bn, tk, nullptr, (flags & 1) != 0)) return RP_ERR_UNSUPPORTED;
    int* o = out96;
    const int head[16] = {H.istr, H.ostr, H.nclass, H.nty, H.ntx, H.tiles_m, H.nplane, H.PH, H.PW, H.NPX, H.a_lbo, H.ntap,
                          H.nkt, H.nkt0, H.tmem_cols, 0};
    for (int i = 0; i < 16; ++i) *o++ = head[i];
    for (int p = 0; p < 4; ++p) { *o++ = H.plane[p].qy; *o++ = H.plane[p].qx; *o++ = H.plane[p].oy; *o++ = H.plane[p].ox; }
    for (int t = 0; t < halo::MAXT; ++t) { *o++ = t < H.ntap ? H.tap[t].cls : 0; *o++ = t < H.ntap ? H.tap[t].a_off : 0; *o++ = t < H.ntap ? H.tap[t].first : 0; }
    for (int c = 0; c < 4; ++c) { *o++ = H.cls_py[c]; *o++ = H.cls_px[c]; *o++ = H.cls_Ha[c]; *o++ = H.cls_Wb[c]; }
    return RP_OK;
}

// Number of halo-kernel launches so far whose input halo was fetched by tiled TMA (cp.async.bulk.tensor).
long long rp_conv_halo_tma_count(void) { return halo::g_tma_launches; }

// Profiling only: read and clear the per-role cycle counters (see g_halo_prof); flags bit 5 enables them.
int rp_conv_halo_prof(unsigned long long* out16) {
    if (!out16) return RP_ERR_INVALID_ARG;
    if (cudaMemcpyFromSymbol(out16, halo::g_halo_prof, sizeof(unsigned long long) * 16) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(halo::g_halo_prof, z, sizeof(z));
    return RP_OK;
}

// Does the tile plan of this layer fit one CTA's shared memory with these flags (the fused split-precision launch, flags bit
// 10, keeps two halos per buffer)?  RP_OK / RP_ERR_UNSUPPORTED; launches nothing.
int rp_conv_halo_fits(const rp_conv_desc* d, int bn, int tk, int flags) {
    halo::HaloArgs H;
    if (!d) return RP_ERR_INVALID_ARG;
    if (!halo::build_halo_args(d, &H, bn, tk, nullptr, (flags & 1) != 0)) return RP_ERR_UNSUPPORTED;
#define RP_HALO_CASE(BN_, TK_, NB_) if (bn == BN_ && tk == TK_) return halo::launch_halo<BN_, TK_, NB_>(H, nullptr, (flags >> 1), nullptr, true);
    RP_HALO_CASE(32, 64, 32) RP_HALO_CASE(64, 64, 12) RP_HALO_CASE(128, 64, 6)
    RP_HALO_CASE(32, 32, 16) RP_HALO_CASE(64, 32, 16) RP_HALO_CASE(128, 32, 8)
    RP_HALO_CASE(32, 16, 16) RP_HALO_CASE(256, 32, 4)
#undef RP_HALO_CASE
    return RP_ERR_UNSUPPORTED;
}

int rp_conv_layer_halo(const rp_conv_desc* d, const void* w_packed, int bn, int tk, int flags, void* stream_) {
    halo::HaloArgs H;
    if (!d || !w_packed || !d->out || !d->src[0].ptr) return RP_ERR_INVALID_ARG;
    if (!halo::build_halo_args(d, &H, bn, tk, nullptr, (flags & 1) != 0)) return RP_ERR_UNSUPPORTED;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
#define RP_HALO_CASE(BN_, TK_, NB_) if (bn == BN_ && tk == TK_) return halo::launch_halo<BN_, TK_, NB_>(H, w_packed, (flags >> 1), stream);
    RP_HALO_CASE(32, 64, 32) RP_HALO_CASE(64, 64, 12) RP_HALO_CASE(128, 64, 6)
    RP_HALO_CASE(32, 32, 16) RP_HALO_CASE(64, 32, 16) RP_HALO_CASE(128, 32, 8)
    RP_HALO_CASE(32, 16, 16) RP_HALO_CASE(256, 32, 4)
#undef RP_HALO_CASE
    return RP_ERR_UNSUPPORTED;
}

}  // extern "C"

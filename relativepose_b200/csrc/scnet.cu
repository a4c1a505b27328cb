// scnet.cu -- SCNet / Resnet18_8s layers for sm_100a (model/mymodel.py of the reference).
//
// The reference runs nn.Conv2d / nn.ConvTranspose2d + BatchNorm2d(track_running_stats=False) + LeakyReLU(0.1)
// through cuDNN/ATen, materialising every torch.cat (mymodel.py:15-39, 266-325).  Here a "conv block" is one
// kernel: while loading its inputs it applies the PRODUCER's batch-norm + LeakyReLU (per scan pair and per
// channel: the reference's BN batch is the 2 images of one forward call), convolves, writes the raw output and
// the partial batch statistics of its own output.  Concatenations are two-source K loops over NHWC views with
// a channel pitch/offset; stride-2 transposed convolutions are decomposed into their 4 sub-pixel classes so no
// multiply-by-zero work is issued.
//
// This file: float32 CUDA-core implicit GEMM (exact-parity path, tolerance 1e-4 max-abs vs the reference
// module) + resize / BN-finalise kernels.  The tcgen05 core for the large layers lives in scnet_tc.cu (round 2).
#include "rp_h16.cuh"
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/rp_b200.h"
#include "scnet_common.cuh"

namespace {
using namespace scnet;

constexpr int BM = 64, BN_ = 64, BK = 16, CT = 256;
// One (group, class, m-tile, n-tile) per CTA.  256 threads: loaders (64 pixels x 4 channel quads / 16 k x 16 n quads),
// compute 16x16 threads x (4 pixels x 4 channels).
__global__ void __launch_bounds__(CT) conv_igemm_f32(const ConvArgs A) {
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN_ + 4];
    __shared__ float red_s[16][BN_];
    __shared__ float red_q[16][BN_];
    const int tid = threadIdx.x;
    const int tile_m = blockIdx.x, tile_n = blockIdx.y;
    const int g = blockIdx.z / A.nclass, ci = blockIdx.z - g * A.nclass;
    const ConvClass& C = A.cls[ci];
    const int HW = C.Ha * C.Wb;
    const int Mc = A.gsz * HW;
    const int tx = tid & 15, ty = tid >> 4;
    const int part_row = (g * A.nclass + ci) * A.tiles_m + tile_m;
    if (tile_m * BM >= Mc) {             // padded tile of a smaller class: contributes zeros to the statistics
        if (A.psum && tid < BN_) {
            int co = tile_n * BN_ + tid;
            if (co < A.Cout) { A.psum[(size_t)part_row * A.Cout + co] = 0.f; A.psq[(size_t)part_row * A.Cout + co] = 0.f; }
        }
        return;
    }
    // loader coordinates (A): pixel pm, channel quad q
    const int pm = tid >> 2, q = tid & 3;
    const int m_l = tile_m * BM + pm;
    const bool mval = m_l < Mc;
    int img_l = 0, a_l = 0, b_l = 0;
    if (mval) { int im = m_l / HW; int rem = m_l - im * HW; img_l = g * A.gsz + im; a_l = rem / C.Wb; b_l = rem - a_l * C.Wb; }
    // loader coordinates (B): k row kb, n quad
    const int kb = tid >> 4, nq = tid & 15;
    const int co_l = tile_n * BN_ + nq * 4;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int t = 0; t < C.ntap; ++t) {
        const Tap tp = C.taps[t];
        const int iy = a_l * A.istr + tp.dy, ix = b_l * A.istr + tp.dx;
        const bool inb = mval && iy >= 0 && iy < A.Hin && ix >= 0 && ix < A.Win;
        const size_t pix = ((size_t)img_l * A.Hin + iy) * A.Win + ix;
        int cbase = 0;
        for (int s = 0; s < A.nsrc; ++s) {
            const rp_conv_src& S = A.src[s];
            for (int c0 = 0; c0 < S.C; c0 += BK) {
                // ---- A tile: 4 channels of one pixel per thread, producer BN + LeakyReLU applied on the fly
                float v[4] = {0.f, 0.f, 0.f, 0.f};
                const int ch = c0 + q * 4;
                if (inb && ch < S.C) {
                    const float* p = S.ptr + pix * S.pitch + S.ch_off + ch;
                    const bool full = (ch + 3 < S.C);
                    if (full && ((((size_t)p) & 15) == 0)) {
                        float4 x = *reinterpret_cast<const float4*>(p);
                        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (ch + j < S.C) v[j] = p[j];
                    }
                    if (S.act) {
                        const float* sc = S.scale + (size_t)g * S.sstride + S.s_off + ch;
                        const float* sh = S.shift + (size_t)g * S.sstride + S.s_off + ch;
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (ch + j < S.C) {
                            float y = fmaf(v[j], sc[j], sh[j]);
                            v[j] = y > 0.f ? y : S.slope * y;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) As[q * 4 + j][pm] = v[j];
                // ---- B tile: packed weights [tap][Cin_total][Cout]
                {
                    float w[4] = {0.f, 0.f, 0.f, 0.f};
                    const int kc = c0 + kb;
                    if (kc < S.C) {
                        const float* wp = A.W + ((size_t)tp.widx * A.Cin_total + cbase + kc) * A.Cout + co_l;
                        if (co_l + 3 < A.Cout && ((((size_t)wp) & 15) == 0)) {
                            float4 x = *reinterpret_cast<const float4*>(wp);
                            w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w;
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) if (co_l + j < A.Cout) w[j] = wp[j];
                        }
                    }
                    *reinterpret_cast<float4*>(&Bs[kb][nq * 4]) = make_float4(w[0], w[1], w[2], w[3]);
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < BK; ++k) {
                    const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                    const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                    acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]); acc[0][2] = fmaf(a.x, b.z, acc[0][2]); acc[0][3] = fmaf(a.x, b.w, acc[0][3]);
                    acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]); acc[1][2] = fmaf(a.y, b.z, acc[1][2]); acc[1][3] = fmaf(a.y, b.w, acc[1][3]);
                    acc[2][0] = fmaf(a.z, b.x, acc[2][0]); acc[2][1] = fmaf(a.z, b.y, acc[2][1]); acc[2][2] = fmaf(a.z, b.z, acc[2][2]); acc[2][3] = fmaf(a.z, b.w, acc[2][3]);
                    acc[3][0] = fmaf(a.w, b.x, acc[3][0]); acc[3][1] = fmaf(a.w, b.y, acc[3][1]); acc[3][2] = fmaf(a.w, b.z, acc[3][2]); acc[3][3] = fmaf(a.w, b.w, acc[3][3]);
                }
                __syncthreads();
            }
            cbase += S.C;
        }
    }

    // ---- epilogue: bias / tanh, store raw output, per-channel partial statistics of this tile
    float ps[4] = {0.f, 0.f, 0.f, 0.f}, pq[4] = {0.f, 0.f, 0.f, 0.f};
    const int co0 = tile_n * BN_ + tx * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = tile_m * BM + ty * 4 + i;
        if (m >= Mc) continue;
        const int im = m / HW; const int rem = m - im * HW; const int a = rem / C.Wb, b = rem - a * C.Wb;
        const int oy = a * A.ostr + C.py, ox = b * A.ostr + C.px;
        float* op = A.out + (((size_t)(g * A.gsz + im) * A.Hout + oy) * A.Wout + ox) * A.out_pitch + A.out_ch_off + co0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (co0 + j >= A.Cout) continue;
            float y = acc[i][j];
            if (A.bias) y += A.bias[co0 + j];
            if (A.tanh_out) y = tanhf(y);
            op[j] = y;
            ps[j] += y; pq[j] += y * y;
        }
    }
    if (A.psum) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { red_s[ty][tx * 4 + j] = ps[j]; red_q[ty][tx * 4 + j] = pq[j]; }
        __syncthreads();
        if (tid < BN_) {
            float s = 0.f, qq = 0.f;
#pragma unroll
            for (int r = 0; r < 16; ++r) { s += red_s[r][tid]; qq += red_q[r][tid]; }
            const int co = tile_n * BN_ + tid;
            if (co < A.Cout) { A.psum[(size_t)part_row * A.Cout + co] = s; A.psq[(size_t)part_row * A.Cout + co] = qq; }
        }
    }
}

// One block per (group, 32 channels): 32 x 32 threads, lanes = channels (coalesced 128-byte rows of the partials),
// warps = row lanes striding over the partial rows; every thread sums its rows in order in float64, the 32 row lanes
// are combined in a fixed order through shared memory -- deterministic, and the 3136-row partials of the 224x224
// layers take 98 independent coalesced loads per thread instead of 98 strided ones per lane.
__global__ void __launch_bounds__(1024) bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int nparts, int Cout,
                                   int count, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ scale, float* __restrict__ shift, int sstride, int s_off) {
    __shared__ double rs[32][33], rq[32][33];
    const int g = blockIdx.y;
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    double s = 0.0, q = 0.0;
    if (c < Cout) {
        const float* ps = psum + (size_t)g * nparts * Cout + c;
        const float* pq = psq + (size_t)g * nparts * Cout + c;
#pragma unroll 4
        for (int r = rl; r < nparts; r += 32) { s += (double)ps[(size_t)r * Cout]; q += (double)pq[(size_t)r * Cout]; }
    }
    rs[rl][cl] = s; rq[rl][cl] = q;
    __syncthreads();
    if (rl == 0 && c < Cout) {
        s = 0.0; q = 0.0;
#pragma unroll
        for (int r = 0; r < 32; ++r) { s += rs[r][cl]; q += rq[r][cl]; }
        const double mean = s / count;
        double var = q / count - mean * mean;          // biased variance, as nn.BatchNorm2d normalises with
        if (var < 0.0) var = 0.0;
        const double sc = (double)gamma[c] / sqrt(var + BN_EPS);
        scale[(size_t)g * sstride + s_off + c] = (float)sc;
        shift[(size_t)g * sstride + s_off + c] = (float)((double)beta[c] - mean * sc);
    }
}

// Two-stage variant for very long partial lists (Resnet18_8s: one BatchNorm batch = the whole call, 12 800+ tile rows):
// stage 1 sums row slices in float64 (grid.z slices), stage 2 combines the slices in order and finalises.
__global__ void __launch_bounds__(1024) bn_partial_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int nparts, int Cout,
                                                          int nsplit, double* __restrict__ scratch) {
    __shared__ double rs[32][33], rq[32][33];
    const int g = blockIdx.y, z = blockIdx.z;
    const int cl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl;
    const int per = (nparts + nsplit - 1) / nsplit, r0 = z * per, r1 = min(nparts, r0 + per);
    double s = 0.0, q = 0.0;
    if (c < Cout) {
        const float* ps = psum + (size_t)g * nparts * Cout + c;
        const float* pq = psq + (size_t)g * nparts * Cout + c;
#pragma unroll 4
        for (int r = r0 + rl; r < r1; r += 32) { s += (double)ps[(size_t)r * Cout]; q += (double)pq[(size_t)r * Cout]; }
    }
    rs[rl][cl] = s; rq[rl][cl] = q;
    __syncthreads();
    if (rl == 0 && c < Cout) {
        s = 0.0; q = 0.0;
#pragma unroll
        for (int r = 0; r < 32; ++r) { s += rs[r][cl]; q += rq[r][cl]; }
        scratch[(((size_t)g * nsplit + z) * Cout + c) * 2] = s;
        scratch[(((size_t)g * nsplit + z) * Cout + c) * 2 + 1] = q;
    }
}

__global__ void bn_combine_kernel(const double* __restrict__ scratch, int nsplit, int Cout, int count, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, float* __restrict__ scale, float* __restrict__ shift, int sstride, int s_off) {
    const int g = blockIdx.y, c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= Cout) return;
    double s = 0.0, q = 0.0;
    for (int z = 0; z < nsplit; ++z) { s += scratch[(((size_t)g * nsplit + z) * Cout + c) * 2]; q += scratch[(((size_t)g * nsplit + z) * Cout + c) * 2 + 1]; }
    const double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double sc = (double)gamma[c] / sqrt(var + BN_EPS);
    scale[(size_t)g * sstride + s_off + c] = (float)sc;
    shift[(size_t)g * sstride + s_off + c] = (float)((double)beta[c] - mean * sc);
}

// im2col of an NHWC float32 image batch into bf16 rows [n, Hout, Wout, Kpad], K index = (ky*k + kx)*C + c, zero padded to
// Kpad (multiple of 32) and outside the image: turns the 7x7/s2 stem of Resnet18_8s (Cin = 7, 49 taps -- too many taps
// and too few channels for the halo kernel's per-tap K chunks) into a 1x1 convolution with K = 352 on tcgen05.
// Space-to-depth, NCHW float32 -> NHWC 16-bit (rp_space_to_depth_h16): one thread per output pixel; its 2x2 input pixels are
// one float2 per (channel, row) -- a warp reads 256 contiguous bytes per (channel, row) -- and it writes Cpad/8 16-byte units.
template <int CT>            // CT > 0: channel count known at compile time (everything stays in registers); 0: runtime C
__global__ void __launch_bounds__(256) space_to_depth_kernel(const float* __restrict__ x, int n, int C_, int H, int W, int Cpad,
                                                             rp_h16* __restrict__ out) {
    const int C = CT > 0 ? CT : C_;
    const int Hs = H >> 1, Ws = W >> 1;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)n * Hs * Ws) return;
    const int sx = (int)(idx % Ws), sy = (int)((idx / Ws) % Hs), im = (int)(idx / ((size_t)Ws * Hs));
    constexpr int NV = CT > 0 ? ((4 * CT + 7) / 8) * 8 : 64;
    rp_h16 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = rp_f_to_h(0.f);
    const float* xb = x + (size_t)im * C * H * W + (size_t)(2 * sy) * W + 2 * sx;
#pragma unroll
    for (int c = 0; c < (CT > 0 ? CT : 16); ++c) {
        if (c < C) {
            const float2 r0 = __ldg(reinterpret_cast<const float2*>(xb + (size_t)c * H * W));
            const float2 r1 = __ldg(reinterpret_cast<const float2*>(xb + (size_t)c * H * W + W));
            v[c] = rp_f_to_h(r0.x); v[C + c] = rp_f_to_h(r0.y); v[2 * C + c] = rp_f_to_h(r1.x); v[3 * C + c] = rp_f_to_h(r1.y);
        }
    }
    uint4* o = reinterpret_cast<uint4*>(out + idx * Cpad);
    const uint4* vv = reinterpret_cast<const uint4*>(v);
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int u = 0; u < 8; ++u) if (u < Cpad / 8) o[u] = u < NV / 8 ? vv[u] : z;
}

constexpr int IC_PX = 32;                          // output pixels (one row segment) per block
constexpr int IC_ROW = 512;                        // floats per staged input row: (IC_PX*s + k - s) * C <= IC_ROW
__global__ void __launch_bounds__(256) im2col_bf16_kernel(const float* __restrict__ x, int n, int H, int W, int C, int k, int s, int p,
                                                          int Hout, int Wout, int Kpad, rp_h16* __restrict__ out) {
    // One block = 32 consecutive output pixels of one output row: the k input rows they read are staged in shared memory
    // once (coalesced), every (pixel, 8-K unit) item is then assembled from shared memory and written as one 16-byte store
    // (a warp writes 512 contiguous bytes).  Without the staging the kernel re-reads its input ~12x through L2.
    __shared__ int s_tab[1024];                    // K index -> offset ky*IC_ROW + kx*C + c into the staged rows, -1 = padding
    __shared__ float patch[7 * IC_ROW];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int kk = tid; kk < Kpad; kk += 256) {
        int e = -1;
        if (kk < k * k * C) { const int tap = kk / C, c = kk - tap * C; const int ky = tap / k, kx = tap - ky * k; e = ky * IC_ROW + kx * C + c; }
        s_tab[kk] = e;
    }
    const int segs = (Wout + IC_PX - 1) / IC_PX;
    const int seg = blockIdx.x % segs;
    const int oy = (blockIdx.x / segs) % Hout;
    const int im = blockIdx.x / (segs * Hout);
    const int ox0 = seg * IC_PX;
    const int pwc = (IC_PX * s + k - s) * C;       // floats of one staged row: contiguous in NHWC memory
    const int iy0 = oy * s - p, g0 = (ox0 * s - p) * C;
    const float* xb = x + (size_t)im * H * W * C;
    for (int r = 0; r < k; ++r) {
        const int iy = iy0 + r;
        const float* row = xb + (size_t)iy * W * C;
        for (int q = tid; q < pwc; q += 256) {
            const int gidx = g0 + q;
            patch[r * IC_ROW + q] = (iy >= 0 && iy < H && gidx >= 0 && gidx < W * C) ? __ldg(row + gidx) : 0.f;
        }
    }
    __syncthreads();
    const int units = Kpad / 8;
    for (int px = warp; px < IC_PX; px += 8) {     // a warp owns a pixel, its lanes the 8-K units: 512 contiguous bytes per store
        if (ox0 + px >= Wout) break;
        const float* pp = patch + px * s * C;
        rp_h16* orow = out + (((size_t)im * Hout + oy) * Wout + ox0 + px) * Kpad;
        for (int u = lane; u < units; u += 32) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { const int e = s_tab[u * 8 + j]; f[j] = e >= 0 ? pp[e] : 0.f; }
            rp_h162 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = rp_f2_to_h2(f[2 * j], f[2 * j + 1]);
            *reinterpret_cast<uint4*>(orow + u * 8) = *reinterpret_cast<uint4*>(v);
        }
    }
}

// ATen upsample_bilinear2d, align_corners=False: src = (dst+0.5)*scale-0.5 clamped at 0, scale = in/out in float.
__device__ __forceinline__ void bilin_coord(int d, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
    float r = scale * ((float)d + 0.5f) - 0.5f;
    if (r < 0.f) r = 0.f;
    i0 = (int)r;
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = r - (float)i0;
    l0 = 1.f - l1;
}

__global__ void scnet_resize_in_kernel(const float* __restrict__ x, int n, int H, int W, float* __restrict__ out) {
    // out [n,224,224,20]; one thread per output pixel
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * 224 * 224) return;
    const int ox = idx % 224, oy = (idx / 224) % 224, im = idx / (224 * 224);
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    bilin_coord(oy, (float)H / 224.f, H, y0, y1, ly0, ly1);
    bilin_coord(ox, (float)W / 224.f, W, x0, x1, lx0, lx1);
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float* p = x + ((size_t)im * 16 + c) * H * W;
        float p00 = p[(size_t)y0 * W + x0], p01 = p[(size_t)y0 * W + x1], p10 = p[(size_t)y1 * W + x0], p11 = p[(size_t)y1 * W + x1];
        v[c] = ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11);
    }
    float* o = out + (size_t)idx * 20;
    // (rgb,mask) (normal,mask) (depth,mask) for the view, then for the warped other view (mymodel.py:264-286)
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[7];
    o[4] = v[3]; o[5] = v[4]; o[6] = v[5]; o[7] = v[7];
    o[8] = v[6]; o[9] = v[7];
    o[10] = v[8]; o[11] = v[9]; o[12] = v[10]; o[13] = v[15];
    o[14] = v[11]; o[15] = v[12]; o[16] = v[13]; o[17] = v[15];
    o[18] = v[14]; o[19] = v[15];
}

// Same resize + regrouping for the tensor-core stem (conv1* on the halo kernel, K chunk 16): out [n,224,224,96]
// bfloat16 = 6 groups x 16 channels, each group = [hi(4) | lo(4) | hi(4) | 0(4)] of its (up to 4) float32 channels,
// x = hi + lo with hi = bf16(x), lo = bf16(x - hi).  Against weights packed as [w_hi | w_hi | w_lo | 0] one K=16 MMA
// per tap computes x_hi*w_hi + x_lo*w_hi + x_hi*w_lo: float32-class accuracy (error ~2^-16) in the K slots a
// 4-channel layer would otherwise pad with zeros.
__global__ void scnet_resize_in_split_kernel(const float* __restrict__ x, int n, int H, int W, rp_h16* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * 224 * 224) return;
    const int ox = idx % 224, oy = (idx / 224) % 224, im = idx / (224 * 224);
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    bilin_coord(oy, (float)H / 224.f, H, y0, y1, ly0, ly1);
    bilin_coord(ox, (float)W / 224.f, W, x0, x1, lx0, lx1);
    float v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) {
        const float* p = x + ((size_t)im * 16 + c) * H * W;
        float p00 = p[(size_t)y0 * W + x0], p01 = p[(size_t)y0 * W + x1], p10 = p[(size_t)y1 * W + x0], p11 = p[(size_t)y1 * W + x1];
        v[c] = ly0 * (lx0 * p00 + lx1 * p01) + ly1 * (lx0 * p10 + lx1 * p11);
    }
    // group -> source channels (mymodel.py:264-286): (rgb,mask) (normal,mask) (depth,mask) of the view, then of the warped view
    const int src[6][4] = {{0, 1, 2, 7}, {3, 4, 5, 7}, {6, 7, -1, -1}, {8, 9, 10, 15}, {11, 12, 13, 15}, {14, 15, -1, -1}};
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)idx * 96);
#pragma unroll
    for (int gq = 0; gq < 6; ++gq) {
        rp_h16 hi[4], lo[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const float f = src[gq][c] >= 0 ? v[src[gq][c]] : 0.f;
            hi[c] = rp_f_to_h(f);
            lo[c] = rp_f_to_h(f - rp_h_to_f(hi[c]));
        }
        rp_h16 u[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) { u[c] = hi[c]; u[4 + c] = lo[c]; u[8 + c] = hi[c]; u[12 + c] = rp_f_to_h(0.f); }
        o[2 * gq] = *reinterpret_cast<uint4*>(&u[0]);
        o[2 * gq + 1] = *reinterpret_cast<uint4*>(&u[8]);
    }
}

// `pitch` floats per pixel in memory; output channel c reads channel cmap[c] (cmap == nullptr: identity) -- the engine
// keeps each head at a 16-byte aligned channel offset so that the head kernels store float4s.
__global__ void scnet_resize_out_kernel(const float* __restrict__ in, int n, int C, int H, int W, float* __restrict__ out,
                                        int pitch, const int* __restrict__ cmap, int out_cs, const int* __restrict__ omap) {
    // in [n,224,224,pitch] NHWC -> out [n,C,H,W]: one thread per output pixel; the four corner pixels are contiguous
    // channel vectors, the per-channel stores are coalesced across the warp (consecutive ox); 4 channels in flight
    __shared__ int s_map[256], s_omap[256];
    for (int i = threadIdx.x; i < C && i < 256; i += blockDim.x) { s_map[i] = cmap ? cmap[i] : i; s_omap[i] = omap ? omap[i] : i; }
    __syncthreads();
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n * H * W;
    if (idx >= total) return;
    const int ox = (int)(idx % W); const int oy = (int)((idx / W) % H); const int im = (int)(idx / ((size_t)W * H));
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    bilin_coord(oy, 224.f / (float)H, 224, y0, y1, ly0, ly1);
    bilin_coord(ox, 224.f / (float)W, 224, x0, x1, lx0, lx1);
    const float* b = in + (size_t)im * 224 * 224 * pitch;
    const float* p00 = b + ((size_t)y0 * 224 + x0) * pitch; const float* p01 = b + ((size_t)y0 * 224 + x1) * pitch;
    const float* p10 = b + ((size_t)y1 * 224 + x0) * pitch; const float* p11 = b + ((size_t)y1 * 224 + x1) * pitch;
    float* o = out + (size_t)im * out_cs * H * W + (size_t)oy * W + ox;      // out_cs: channels per image of the output tensor (>= C)
    int c = 0;
    for (; c + 4 <= C; c += 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = s_map[c + u];
            v[u] = ly0 * (lx0 * __ldg(p00 + s) + lx1 * __ldg(p01 + s)) + ly1 * (lx0 * __ldg(p10 + s) + lx1 * __ldg(p11 + s));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) o[(size_t)s_omap[c + u] * H * W] = v[u];
    }
    for (; c < C; ++c) {
        const int s = s_map[c];
        o[(size_t)s_omap[c] * H * W] = ly0 * (lx0 * p00[s] + lx1 * p01[s]) + ly1 * (lx0 * p10[s] + lx1 * p11[s]);
    }
}

// Direct 3x3/s1/p1 convolution for the tiny-Cin encoder stems (conv1rgb/conv1n: Cin=4, conv1d: Cin=2 -> 32 channels,
// mymodel.py:151,155,159): one thread per output pixel x 32 channels, weights + nothing else in smem.  The implicit
// GEMM kernel would spend 75-88 % of every 16-wide K tile on padding here.
template <int CIN>
__global__ void __launch_bounds__(128) conv3x3_small_cin(const ConvArgs A) {
    __shared__ float Ws[9 * CIN * 32];
    __shared__ float red_s[4][32], red_q[4][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 9 * CIN * 32; i += 128) Ws[i] = A.W[i];          // packed [tap][ci][co], Cout == 32
    __syncthreads();
    const int g = blockIdx.y;
    const int HW = A.Hout * A.Wout;
    const int m = blockIdx.x * 128 + tid;                                   // pixel within the pair (2 images)
    const bool val = m < 2 * HW;
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    int im = 0, oy = 0, ox = 0;
    if (val) {
        im = m / HW; const int rem = m - im * HW; oy = rem / A.Wout; ox = rem - oy * A.Wout;
        const rp_conv_src& S = A.src[0];
        const float* base = S.ptr + (size_t)(g * 2 + im) * A.Hin * A.Win * S.pitch + S.ch_off;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy + ky - 1;
            if (iy < 0 || iy >= A.Hin) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = ox + kx - 1;
                if (ix < 0 || ix >= A.Win) continue;
                const float* p = base + ((size_t)iy * A.Win + ix) * S.pitch;
                float x[CIN];
#pragma unroll
                for (int c = 0; c < CIN; ++c) x[c] = p[c];
                const float* w = Ws + (ky * 3 + kx) * CIN * 32;
#pragma unroll
                for (int c = 0; c < CIN; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = fmaf(x[c], w[c * 32 + j], acc[j]);
            }
        }
        const size_t e = (((size_t)(g * 2 + im) * A.Hout + oy) * A.Wout + ox) * A.out_pitch + A.out_ch_off;
        if (A.out_bf16) {          // bfloat16 storage: round first so the statistics describe the stored tensor
            rp_h16* op = reinterpret_cast<rp_h16*>(A.out) + e;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                rp_h162 p[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    p[q] = rp_f2_to_h2(acc[j + 2 * q], acc[j + 2 * q + 1]);
                    const float2 f = rp_h2_to_f2(p[q]);
                    acc[j + 2 * q] = f.x; acc[j + 2 * q + 1] = f.y;
                }
                *reinterpret_cast<uint4*>(op + j) = *reinterpret_cast<uint4*>(p);
            }
        } else {
            float* op = A.out + e;
#pragma unroll
            for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(op + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
        }
    }
    if (A.psum) {       // per-channel sums over the 128 pixels of this CTA: butterfly per warp, 4 warps combined in order
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            float s1 = acc[j], s2 = acc[j] * acc[j];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
            if (lane == 0) { red_s[warp][j] = s1; red_q[warp][j] = s2; }
        }
        __syncthreads();
        if (tid < 32) {
            const size_t row = (size_t)g * gridDim.x + blockIdx.x;
            A.psum[row * 32 + tid] = ((red_s[0][tid] + red_s[1][tid]) + red_s[2][tid]) + red_s[3][tid];
            A.psq[row * 32 + tid] = ((red_q[0][tid] + red_q[1][tid]) + red_q[2][tid]) + red_q[3][tid];
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Resnet18_8s extras (mymodel.py:82-122): all float32 NHWC, BN scale/shift per (group, channel).
__global__ void bn_relu_maxpool_kernel(const float* __restrict__ in, int n, int H, int W, int C, int gsz,
                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                       float* __restrict__ out, int Ho, int Wo) {
    // 4 channels per thread (C % 4 == 0): float4 loads / stores
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int C4 = C / 4;
    const size_t total = (size_t)n * Ho * Wo * C4;
    if (i4 >= total) return;
    const int c = (int)(i4 % C4) * 4; const int ox = (int)((i4 / C4) % Wo); const int oy = (int)((i4 / ((size_t)C4 * Wo)) % Ho);
    const int im = (int)(i4 / ((size_t)C4 * Wo * Ho));
    const int g = im / gsz;
    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
    if (scale) sc = *reinterpret_cast<const float4*>(scale + (size_t)g * C + c);
    if (shift) sh = *reinterpret_cast<const float4*>(shift + (size_t)g * C + c);
    float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            if (ix < 0 || ix >= W) continue;
            const float4 v = *reinterpret_cast<const float4*>(in + (((size_t)im * H + iy) * W + ix) * C + c);
            m.x = fmaxf(m.x, fmaxf(fmaf(v.x, sc.x, sh.x), 0.f)); m.y = fmaxf(m.y, fmaxf(fmaf(v.y, sc.y, sh.y), 0.f));
            m.z = fmaxf(m.z, fmaxf(fmaf(v.z, sc.z, sh.z), 0.f)); m.w = fmaxf(m.w, fmaxf(fmaf(v.w, sc.w, sh.w), 0.f));
        }
    }
    *reinterpret_cast<float4*>(out + i4 * 4) = m;
}

__global__ void bn_add_relu_kernel(const float* __restrict__ a, const float* __restrict__ sa, const float* __restrict__ ha,
                                   const float* __restrict__ b, const float* __restrict__ sb, const float* __restrict__ hb,
                                   float* __restrict__ out, int n, int HW, int C, int gsz) {
    // 4 channels per thread (C % 4 == 0): float4 loads / stores, HBM-bound
    const size_t i4 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total4 = (size_t)n * HW * C / 4;
    if (i4 >= total4) return;
    const size_t idx = i4 * 4;
    const int c = (int)(idx % C);
    const int g = (int)(idx / ((size_t)C * HW)) / gsz;
    const float4 av = *reinterpret_cast<const float4*>(a + idx), bv = *reinterpret_cast<const float4*>(b + idx);
    const float4 s1 = *reinterpret_cast<const float4*>(sa + (size_t)g * C + c), h1 = *reinterpret_cast<const float4*>(ha + (size_t)g * C + c);
    float4 y = bv;
    if (sb) {
        const float4 s2 = *reinterpret_cast<const float4*>(sb + (size_t)g * C + c), h2 = *reinterpret_cast<const float4*>(hb + (size_t)g * C + c);
        y.x = fmaf(bv.x, s2.x, h2.x); y.y = fmaf(bv.y, s2.y, h2.y); y.z = fmaf(bv.z, s2.z, h2.z); y.w = fmaf(bv.w, s2.w, h2.w);
    }
    float4 o;
    o.x = fmaxf(fmaf(av.x, s1.x, h1.x) + y.x, 0.f); o.y = fmaxf(fmaf(av.y, s1.y, h1.y) + y.y, 0.f);
    o.z = fmaxf(fmaf(av.z, s1.z, h1.z) + y.z, 0.f); o.w = fmaxf(fmaf(av.w, s1.w, h1.w) + y.w, 0.f);
    *reinterpret_cast<float4*>(out + idx) = o;
}

__global__ void resize_nhwc_kernel(const float* __restrict__ src, int n, int Hs, int Ws, int C, float* __restrict__ dst,
                                   int Hd, int Wd, int accumulate) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n * Hd * Wd * C;
    if (idx >= total) return;
    const int c = (int)(idx % C); const int ox = (int)((idx / C) % Wd); const int oy = (int)((idx / ((size_t)C * Wd)) % Hd);
    const int im = (int)(idx / ((size_t)C * Wd * Hd));
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    bilin_coord(oy, (float)Hs / (float)Hd, Hs, y0, y1, ly0, ly1);
    bilin_coord(ox, (float)Ws / (float)Wd, Ws, x0, x1, lx0, lx1);
    const float* p = src + (size_t)im * Hs * Ws * C + c;
    float v = ly0 * (lx0 * p[((size_t)y0 * Ws + x0) * C] + lx1 * p[((size_t)y0 * Ws + x1) * C]) +
              ly1 * (lx0 * p[((size_t)y1 * Ws + x0) * C] + lx1 * p[((size_t)y1 * Ws + x1) * C]);
    dst[idx] = accumulate ? dst[idx] + v : v;
}

__global__ void resize_to_nchw_kernel(const float* __restrict__ src, int n, int Hs, int Ws, int C, float* __restrict__ out,
                                      int H, int W, int tanh_out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n * H * W;
    if (idx >= total) return;
    const int ox = (int)(idx % W); const int oy = (int)((idx / W) % H); const int im = (int)(idx / ((size_t)W * H));
    int y0, y1, x0, x1; float ly0, ly1, lx0, lx1;
    bilin_coord(oy, (float)Hs / (float)H, Hs, y0, y1, ly0, ly1);
    bilin_coord(ox, (float)Ws / (float)W, Ws, x0, x1, lx0, lx1);
    const float* b = src + (size_t)im * Hs * Ws * C;
    const float* p00 = b + ((size_t)y0 * Ws + x0) * C; const float* p01 = b + ((size_t)y0 * Ws + x1) * C;
    const float* p10 = b + ((size_t)y1 * Ws + x0) * C; const float* p11 = b + ((size_t)y1 * Ws + x1) * C;
    float* o = out + (size_t)im * C * H * W + (size_t)oy * W + ox;
    for (int c = 0; c < C; ++c) {
        float v = ly0 * (lx0 * p00[c] + lx1 * p01[c]) + ly1 * (lx0 * p10[c] + lx1 * p11[c]);
        float y = v;
        if (tanh_out == 1) y = tanhf(v);
        else if (tanh_out == 2) asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(v));      // tensor-core mode: error ~5e-4, far below bf16
        o[(size_t)c * H * W] = y;
    }
}

// Up-sampling NHWC [n,Hs,Ws,pitch] -> NCHW [n,C,H,W] (ATen bilinear, align_corners=False) with the two source rows of an
// output row segment staged in shared memory: the per-pixel version gathers 4 corner vectors per thread, ~11 distinct
// 128-byte lines per warp load, and is LSU-bound at ~1.7 TB/s of output.  One block = 64 consecutive output pixels of
// one row; global loads are contiguous runs, output stores are 256-byte runs per channel.
constexpr int RS_PX = 64, RS_MAXSRC = 40, RS_MAXFL = 2 * RS_MAXSRC * 72;
__global__ void __launch_bounds__(256) resize_up_nchw_kernel(const float* __restrict__ src, int n, int Hs, int Ws, int pitch,
                                                             const int* __restrict__ cmap, int C, float* __restrict__ out, int H, int W,
                                                             int tanh_out, int out_cs, const int* __restrict__ omap) {
    __shared__ __align__(16) float rows[RS_MAXFL];
    __shared__ int s_x0[RS_PX], s_x1[RS_PX], s_map[256], s_omap[256];
    __shared__ float s_l0[RS_PX], s_l1[RS_PX];
    const int tid = threadIdx.x;
    const int segs = (W + RS_PX - 1) / RS_PX;
    const int seg = blockIdx.x % segs, oy = (blockIdx.x / segs) % H, im = blockIdx.x / (segs * H);
    const int ox0 = seg * RS_PX;
    int y0, y1; float ly0, ly1;
    bilin_coord(oy, (float)Hs / (float)H, Hs, y0, y1, ly0, ly1);
    if (tid < RS_PX) {
        int x0 = 0, x1 = 0; float l0 = 0.f, l1 = 0.f;
        if (ox0 + tid < W) bilin_coord(ox0 + tid, (float)Ws / (float)W, Ws, x0, x1, l0, l1);
        s_x0[tid] = x0; s_x1[tid] = x1; s_l0[tid] = l0; s_l1[tid] = l1;
    }
    for (int i = tid; i < C; i += 256) { s_map[i] = cmap ? cmap[i] : i; s_omap[i] = omap ? omap[i] : i; }   // item i: source / output channel
    __syncthreads();
    const int last = min(W - 1, ox0 + RS_PX - 1) - ox0;
    const int xs0 = s_x0[0], nsrc = s_x1[last] - xs0 + 1;       // source columns of the segment (<= RS_MAXSRC, checked on the host)
    const float* b = src + (size_t)im * Hs * Ws * pitch;
    const int run = nsrc * pitch;
    const float* g0 = b + ((size_t)y0 * Ws + xs0) * pitch;
    const float* g1 = b + ((size_t)y1 * Ws + xs0) * pitch;
    if ((pitch & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {      // rows start 16-byte aligned: float4 staging
        const int run4 = run >> 2;
        float4* d0 = reinterpret_cast<float4*>(rows); float4* d1 = reinterpret_cast<float4*>(rows + run);
        for (int i = tid; i < run4; i += 256) {
            d0[i] = __ldg(reinterpret_cast<const float4*>(g0) + i);
            d1[i] = __ldg(reinterpret_cast<const float4*>(g1) + i);
        }
    } else {
        for (int i = tid; i < run; i += 256) {
            rows[i] = __ldg(g0 + i);
            rows[run + i] = __ldg(g1 + i);
        }
    }
    __syncthreads();
    const size_t plane = (size_t)H * W;
    float* o = out + (size_t)im * out_cs * plane + (size_t)oy * W + ox0;     // out_cs: channels per image of the output tensor (>= C)
    // a thread keeps ONE output pixel (its column offsets and weights stay in registers) and walks the channels: a warp
    // writes 32 consecutive pixels of one channel, and the inner loop is 4 shared-memory reads + the channel map per output
    const int px = tid & (RS_PX - 1);
    if (ox0 + px >= W) return;
    const int a0 = (s_x0[px] - xs0) * pitch, a1 = (s_x1[px] - xs0) * pitch;
    const float lx0 = s_l0[px], lx1 = s_l1[px];
    const float* r0a = rows + a0; const float* r0b = rows + a1;
    const float* r1a = rows + run + a0; const float* r1b = rows + run + a1;
    o += px;
#pragma unroll 4
    for (int c = tid / RS_PX; c < C; c += 256 / RS_PX) {
        const int sc = s_map[c];
        const float v = ly0 * (lx0 * r0a[sc] + lx1 * r0b[sc]) + ly1 * (lx0 * r1a[sc] + lx1 * r1b[sc]);
        float y = v;
        if (tanh_out == 1) y = tanhf(v);
        else if (tanh_out == 2) asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(v));
        __stcs(o + (size_t)s_omap[c] * plane, y);
    }
}

// staged version applicable: up-sampling in x by enough that a 64-pixel segment reads <= RS_MAXSRC source columns
inline bool resize_up_ok(int Ws, int W, int pitch, int C) {
    if (W < Ws || C > 256 || pitch > 72) return false;
    const int nsrc = (int)((double)RS_PX * Ws / W) + 3;
    return nsrc <= RS_MAXSRC;
}

// rputil.interpolate (RPModule/rputil.py:43-58): bilinear gather of C-channel descriptors at K normalised points,
// x = px*(W-1), y = py*(H-1), floor-based weights, float32 with the reference's operation order.
__global__ void interpolate_kernel(const float* __restrict__ feat, int C, int H, int W, const float* __restrict__ pt, int K,
                                   float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= C * K) return;
    const int k = idx % K, c = idx / K;
    const float x = pt[2 * k] * (float)(W - 1), y = pt[2 * k + 1] * (float)(H - 1);
    const float x0 = floorf(x), y0 = floorf(y);
    int ix = (int)x0, iy = (int)y0;
    ix = ix < 0 ? 0 : (ix > W - 2 ? W - 2 : ix);
    iy = iy < 0 ? 0 : (iy > H - 2 ? H - 2 : iy);
    const float* f = feat + (size_t)c * H * W;
    const float v00 = f[(size_t)iy * W + ix], v10 = f[(size_t)(iy + 1) * W + ix];
    const float v01 = f[(size_t)iy * W + ix + 1], v11 = f[(size_t)(iy + 1) * W + ix + 1];
    const float wx0 = __fsub_rn(__fadd_rn(x0, 1.f), x), wy0 = __fsub_rn(__fadd_rn(y0, 1.f), y);
    const float wx1 = __fsub_rn(x, x0), wy1 = __fsub_rn(y, y0);
    float r = __fmul_rn(__fmul_rn(v00, wx0), wy0);
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(v10, wx0), wy1));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(v01, wx1), wy0));
    r = __fadd_rn(r, __fmul_rn(__fmul_rn(v11, wx1), wy1));
    out[idx] = r;
}

// 1x1 output heads (deconv1rgb/n/d/s/f, mymodel.py:188-228: Conv2d(k=1) + bias [+ tanh], no BatchNorm): a K <= 128 by
// Cout <= 32 matrix per pixel.  HBM-bound (the 3-channel heads) or FMA-bound (the 21/32-channel heads) -- far too
// little work per 128-pixel tile to amortise a tensor-core CTA's setup, so: CUDA cores, PIX pixels per thread, the
// zero-padded [K][CP] weights broadcast from shared memory, producer BatchNorm + LeakyReLU applied while loading
// (float32 or bfloat16 storage).
template <int CP, int PIX, int CH>      // CP: padded Cout, PIX: pixels per thread, CH: channels loaded per step (8 or 32)
__global__ void __launch_bounds__(128) conv1x1_head_kernel(const ConvArgs A) {
    __shared__ __align__(16) float Ws[128 * CP];
    __shared__ float s_sc[128], s_sh[128];
    const int tid = threadIdx.x, g = blockIdx.y;
    for (int i = tid; i < A.Cin_total * CP; i += 128) { const int k = i / CP, j = i - k * CP; Ws[i] = j < A.Cout ? A.W[(size_t)k * A.Cout + j] : 0.f; }
    {
        int cb = 0;
        for (int si = 0; si < A.nsrc; ++si) {
            const rp_conv_src& S = A.src[si];
            for (int c = tid; c < S.C; c += 128) {
                s_sc[cb + c] = S.act ? S.scale[(size_t)g * S.sstride + S.s_off + c] : 1.f;
                s_sh[cb + c] = S.act ? S.shift[(size_t)g * S.sstride + S.s_off + c] : 0.f;
            }
            cb += S.C;
        }
    }
    __syncthreads();
    const int npx = A.gsz * A.Hout * A.Wout;
    constexpr int NU = CH / 8;                                   // 8-channel units per step
    int px[PIX]; bool val[PIX];
    float acc[PIX][CP];
#pragma unroll
    for (int u = 0; u < PIX; ++u) {
        px[u] = blockIdx.x * (128 * PIX) + u * 128 + tid;
        val[u] = px[u] < npx;
#pragma unroll
        for (int j = 0; j < CP; ++j) acc[u][j] = 0.f;
    }
    int cb = 0;
    for (int si = 0; si < A.nsrc; ++si) {
        const rp_conv_src& S = A.src[si];
        const bool h16 = S.dtype == 1;
        for (int c0 = 0; c0 < S.C; c0 += CH) {
            uint4 raw[PIX][NU][2];                               // every load of the step is issued before any is used
#pragma unroll
            for (int u = 0; u < PIX; ++u) {
                if (!val[u]) continue;
                const size_t e = ((size_t)g * npx + px[u]) * S.pitch + S.ch_off + c0;
#pragma unroll
                for (int w = 0; w < NU; ++w) {
                    if (h16) {
                        raw[u][w][0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const rp_h16*>(S.ptr) + e + 8 * w);
                    } else {
                        raw[u][w][0] = *reinterpret_cast<const uint4*>(S.ptr + e + 8 * w);
                        raw[u][w][1] = *reinterpret_cast<const uint4*>(S.ptr + e + 8 * w + 4);
                    }
                }
            }
#pragma unroll
            for (int w = 0; w < NU; ++w) {
                float v[PIX][8];
#pragma unroll
                for (int u = 0; u < PIX; ++u) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[u][q] = 0.f;
                    if (!val[u]) continue;
                    if (h16) {
                        const rp_h162* hp = reinterpret_cast<const rp_h162*>(&raw[u][w][0]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) { const float2 f = rp_h2_to_f2(hp[q]); v[u][2 * q] = f.x; v[u][2 * q + 1] = f.y; }
                    } else {
                        v[u][0] = __uint_as_float(raw[u][w][0].x); v[u][1] = __uint_as_float(raw[u][w][0].y);
                        v[u][2] = __uint_as_float(raw[u][w][0].z); v[u][3] = __uint_as_float(raw[u][w][0].w);
                        v[u][4] = __uint_as_float(raw[u][w][1].x); v[u][5] = __uint_as_float(raw[u][w][1].y);
                        v[u][6] = __uint_as_float(raw[u][w][1].z); v[u][7] = __uint_as_float(raw[u][w][1].w);
                    }
                    if (S.act) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float z = fmaf(v[u][q], s_sc[cb + c0 + 8 * w + q], s_sh[cb + c0 + 8 * w + q]);
                            v[u][q] = z > 0.f ? z : S.slope * z;
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float* wr = Ws + (cb + c0 + 8 * w + q) * CP;
#pragma unroll
                    for (int j = 0; j < CP; j += 4) {
                        const float4 w4 = *reinterpret_cast<const float4*>(wr + j);
#pragma unroll
                        for (int u = 0; u < PIX; ++u) {
                            acc[u][j] = fmaf(v[u][q], w4.x, acc[u][j]); acc[u][j + 1] = fmaf(v[u][q], w4.y, acc[u][j + 1]);
                            acc[u][j + 2] = fmaf(v[u][q], w4.z, acc[u][j + 2]); acc[u][j + 3] = fmaf(v[u][q], w4.w, acc[u][j + 3]);
                        }
                    }
                }
            }
        }
        cb += S.C;
    }
#pragma unroll
    for (int u = 0; u < PIX; ++u) {
        if (!val[u]) continue;
        float* op = A.out + ((size_t)g * npx + px[u]) * A.out_pitch + A.out_ch_off;
#pragma unroll
        for (int j = 0; j < CP; ++j) {
            if (j < A.Cout) { float y = acc[u][j] + (A.bias ? A.bias[j] : 0.f); op[j] = A.tanh_out ? tanhf(y) : y; }
        }
    }
}

inline bool head_eligible(const rp_conv_desc* d) {
    if (d->transposed || d->k != 1 || d->s != 1 || d->p != 0 || d->psum || d->Cout > 32 || d->out_dtype != 0) return false;
    if (d->Hin != d->Hout || d->Win != d->Wout) return false;
    int K = 0;
    for (int i = 0; i < d->nsrc; ++i) {
        const int al = d->src[i].dtype == 1 ? 8 : 4;
        if (d->src[i].C % 8 || d->src[i].pitch % al || d->src[i].ch_off % al) return false;
        K += d->src[i].C;
    }
    return K <= 128;
}

inline bool small_cin_eligible(const rp_conv_desc* d) {
    return d->nsrc == 1 && !d->transposed && d->k == 3 && d->s == 1 && d->p == 1 && d->Cout == 32 && !d->bias &&
           !d->tanh_out && d->src[0].act == 0 && (d->src[0].C == 4 || d->src[0].C == 2) &&
           d->src[0].dtype == 0 && (d->out_pitch % (d->out_dtype == 1 ? 8 : 4) == 0) && (d->out_ch_off % (d->out_dtype == 1 ? 8 : 4) == 0) && d->Hin == d->Hout && d->Win == d->Wout && (d->imgs_per_group == 0 || d->imgs_per_group == 2);
}

}  // namespace
namespace scnet { long long g_conv_launches = 0; }
namespace {

}  // namespace

extern "C" {

int rp_conv_nparts(const rp_conv_desc* d, int* nparts) {
    ConvArgs A;
    if (!nparts || !build_args(d, &A, BM)) return RP_ERR_INVALID_ARG;
    if (small_cin_eligible(d)) { *nparts = (2 * d->Hout * d->Wout + 127) / 128; return RP_OK; }
    *nparts = A.nclass * A.tiles_m;
    return RP_OK;
}

int rp_conv_layer(const rp_conv_desc* d, void* stream_) {
    ConvArgs A;
    if (!build_args(d, &A, BM)) return RP_ERR_INVALID_ARG;
    if (!d->W || !d->out || !d->src[0].ptr) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (small_cin_eligible(d)) {
        dim3 grid((2 * d->Hout * d->Wout + 127) / 128, A.G);
        if (d->src[0].C == 4) conv3x3_small_cin<4><<<grid, 128, 0, stream>>>(A);
        else conv3x3_small_cin<2><<<grid, 128, 0, stream>>>(A);
        ++g_conv_launches;
        return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
    }
    if (head_eligible(d)) {
        const int npx = A.gsz * A.Hout * A.Wout;
        bool c32 = true;
        for (int i = 0; i < d->nsrc; ++i) c32 = c32 && (d->src[i].C % 32 == 0);
        if (A.Cout <= 4 && c32) { dim3 grid((npx + 255) / 256, A.G); conv1x1_head_kernel<4, 2, 32><<<grid, 128, 0, stream>>>(A); }
        else if (A.Cout <= 4) { dim3 grid((npx + 511) / 512, A.G); conv1x1_head_kernel<4, 4, 8><<<grid, 128, 0, stream>>>(A); }
        else if (A.Cout <= 16) { dim3 grid((npx + 255) / 256, A.G); conv1x1_head_kernel<16, 2, 8><<<grid, 128, 0, stream>>>(A); }
        else if (A.Cout <= 24) { dim3 grid((npx + 255) / 256, A.G); conv1x1_head_kernel<24, 2, 8><<<grid, 128, 0, stream>>>(A); }
        else { dim3 grid((npx + 255) / 256, A.G); conv1x1_head_kernel<32, 2, 8><<<grid, 128, 0, stream>>>(A); }
        ++g_conv_launches;
        return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
    }
    if (d->out_dtype != 0) return RP_ERR_UNSUPPORTED;          // the float32 implicit GEMM reads and writes float32 only
    for (int i = 0; i < d->nsrc; ++i) if (d->src[i].dtype != 0) return RP_ERR_UNSUPPORTED;
    dim3 grid(A.tiles_m, (A.Cout + BN_ - 1) / BN_, A.G * A.nclass);
    conv_igemm_f32<<<grid, CT, 0, stream>>>(A);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_bn_finalize(const float* psum, const float* psq, int G, int nparts, int Cout, int count,
                   const float* gamma, const float* beta, float* scale, float* shift, int sstride, int s_off,
                   void* stream_) {
    if (!psum || !psq || !gamma || !beta || !scale || !shift || G < 1 || nparts < 1 || Cout < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    dim3 grid((Cout + 31) / 32, G);                   // 32 channels x 32 row lanes per block
    bn_finalize_kernel<<<grid, 1024, 0, stream>>>(psum, psq, nparts, Cout, count, gamma, beta, scale, shift, sstride, s_off);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_bn_finalize_split(const float* psum, const float* psq, int G, int nparts, int Cout, int count,
                         const float* gamma, const float* beta, float* scale, float* shift, int sstride, int s_off,
                         int nsplit, double* scratch, void* stream_) {
    if (!psum || !psq || !gamma || !beta || !scale || !shift || !scratch || G < 1 || nparts < 1 || Cout < 1 || nsplit < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    dim3 g1((Cout + 31) / 32, G, nsplit);
    bn_partial_kernel<<<g1, 1024, 0, stream>>>(psum, psq, nparts, Cout, nsplit, scratch);
    dim3 g2((Cout + 127) / 128, G);
    bn_combine_kernel<<<g2, 128, 0, stream>>>(scratch, nsplit, Cout, count, gamma, beta, scale, shift, sstride, s_off);
    g_conv_launches += 2;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_im2col_bf16(const float* x, int n, int H, int W, int C, int k, int s, int p, int Hout, int Wout, int Kpad, void* out, void* stream_) {
    if (!x || !out || n < 1 || C < 1 || C > 1023 || k < 1 || k > 31 || s < 1 || Kpad < k * k * C || (Kpad % 8) || Kpad > 1024) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (k > 7 || (IC_PX * s + k - s) * C > IC_ROW) return RP_ERR_UNSUPPORTED;
    const size_t blocks = (size_t)n * Hout * ((Wout + IC_PX - 1) / IC_PX);
    im2col_bf16_kernel<<<(unsigned)blocks, 256, 0, stream>>>(x, n, H, W, C, k, s, p, Hout, Wout, Kpad, static_cast<rp_h16*>(out));
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_space_to_depth_h16(const float* x, int n, int C, int H, int W, int Cpad, void* out, void* stream_) {
    if (!x || !out || n < 1 || C < 1) return RP_ERR_INVALID_ARG;
    if ((H & 1) || (W & 1) || C > 16 || 4 * C > Cpad || (Cpad & 7) || Cpad > 64) return RP_ERR_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(x) & 7) != 0) return RP_ERR_INVALID_ARG;                 // float2 loads
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const size_t total = (size_t)n * (H / 2) * (W / 2);
    const unsigned nb = (unsigned)((total + 255) / 256);
    if (C == 7) space_to_depth_kernel<7><<<nb, 256, 0, stream>>>(x, n, C, H, W, Cpad, static_cast<rp_h16*>(out));
    else space_to_depth_kernel<0><<<nb, 256, 0, stream>>>(x, n, C, H, W, Cpad, static_cast<rp_h16*>(out));
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_scnet_resize_in(const float* x, int n, int H, int W, float* out, void* stream_) {
    if (!x || !out || n < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int total = n * 224 * 224;
    scnet_resize_in_kernel<<<(total + 255) / 256, 256, 0, stream>>>(x, n, H, W, out);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_scnet_resize_in_split(const float* x, int n, int H, int W, void* out, void* stream_) {
    if (!x || !out || n < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int total = n * 224 * 224;
    scnet_resize_in_split_kernel<<<(total + 127) / 128, 128, 0, stream>>>(x, n, H, W, static_cast<rp_h16*>(out));
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_scnet_resize_out(const float* in, int n, int C, int H, int W, float* out, void* stream_) {
    if (!in || !out || n < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    size_t total = (size_t)n * H * W;
    if (resize_up_ok(224, W, C, C) && H >= 1) resize_up_nchw_kernel<<<(unsigned)((size_t)n * H * ((W + RS_PX - 1) / RS_PX)), 256, 0, stream>>>(in, n, 224, 224, C, nullptr, C, out, H, W, 0, C, nullptr);
    else scnet_resize_out_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, n, C, H, W, out, C, nullptr, C, nullptr);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_scnet_resize_out_map(const float* in, int n, int pitch, const int* cmap, int C, int H, int W, float* out, void* stream_) {
    if (!in || !out || !cmap || n < 1 || pitch < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    size_t total = (size_t)n * H * W;
    if (resize_up_ok(224, W, pitch, C)) resize_up_nchw_kernel<<<(unsigned)((size_t)n * H * ((W + RS_PX - 1) / RS_PX)), 256, 0, stream>>>(in, n, 224, 224, pitch, cmap, C, out, H, W, 0, C, nullptr);
    else scnet_resize_out_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, n, C, H, W, out, pitch, cmap, C, nullptr);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_scnet_resize_out_sub(const float* in, int n, int pitch, const int* cmap, const int* omap, int C, int H, int W, float* out,
                            int out_channels, void* stream_) {
    if (!in || !out || !cmap || !omap || n < 1 || pitch < 1 || C < 1 || C > 256 || C > out_channels) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    size_t total = (size_t)n * H * W;
    if (resize_up_ok(224, W, pitch, C)) resize_up_nchw_kernel<<<(unsigned)((size_t)n * H * ((W + RS_PX - 1) / RS_PX)), 256, 0, stream>>>(in, n, 224, 224, pitch, cmap, C, out, H, W, 0, out_channels, omap);
    else scnet_resize_out_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, n, C, H, W, out, pitch, cmap, out_channels, omap);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_bn_relu_maxpool(const float* in, int n, int H, int W, int C, int imgs_per_group,
                       const float* scale, const float* shift, float* out, int Ho, int Wo, void* stream_) {
    if (!in || !out || n < 1 || imgs_per_group < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (C % 4) return RP_ERR_UNSUPPORTED;
    size_t total = (size_t)n * Ho * Wo * C / 4;
    bn_relu_maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(in, n, H, W, C, imgs_per_group, scale, shift, out, Ho, Wo);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_bn_add_relu(const float* a, const float* sa, const float* ha, const float* b, const float* sb, const float* hb,
                   float* out, int n, int HW, int C, int imgs_per_group, void* stream_) {
    if (!a || !sa || !ha || !b || !out || n < 1 || imgs_per_group < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (C % 4) return RP_ERR_UNSUPPORTED;
    size_t total = (size_t)n * HW * C / 4;
    bn_add_relu_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(a, sa, ha, b, sb, hb, out, n, HW, C, imgs_per_group);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_resize_nhwc(const float* src, int n, int Hs, int Ws, int C, float* dst, int Hd, int Wd, int accumulate, void* stream_) {
    if (!src || !dst || n < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    size_t total = (size_t)n * Hd * Wd * C;
    resize_nhwc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, n, Hs, Ws, C, dst, Hd, Wd, accumulate);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_resize_to_nchw(const float* src, int n, int Hs, int Ws, int C, float* out, int H, int W, int tanh_out, void* stream_) {
    if (!src || !out || n < 1) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    size_t total = (size_t)n * H * W;
    if (resize_up_ok(Ws, W, C, C)) resize_up_nchw_kernel<<<(unsigned)((size_t)n * H * ((W + RS_PX - 1) / RS_PX)), 256, 0, stream>>>(src, n, Hs, Ws, C, nullptr, C, out, H, W, tanh_out, C, nullptr);
    else resize_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(src, n, Hs, Ws, C, out, H, W, tanh_out);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_interpolate(const float* feat, int C, int H, int W, const float* pt, int K, float* out, void* stream_) {
    if (!feat || !pt || !out || C < 1 || K < 0 || H < 2 || W < 2) return RP_ERR_INVALID_ARG;
    if (K == 0) return RP_OK;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    interpolate_kernel<<<(C * K + 255) / 256, 256, 0, stream>>>(feat, C, H, W, pt, K, out);
    ++g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int64_t rp_conv_launch_count(void) { return g_conv_launches; }

}  // extern "C"

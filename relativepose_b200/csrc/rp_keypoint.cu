// rp_keypoint.cu -- keypoint augmentation, sm_100a.
//
// Reference (RPModule/rputil.py): for ~30-100 selected keypoints of one scan, the squared descriptor distance to EVERY
// pixel of the other scan's 32-channel feature map is formed on the GPU as a [C, n, H*W] broadcast temporary
// (`(f[:, sel].unsqueeze(2) - feat.view(C,1,-1)).pow(2).sum(0)`, :187,189,209), copied to the host, and `Sampling` (:355-371)
// runs K=2 rounds of { argmax of exp(-d/2), suppress a 15-pixel window with the map's minimum } in a Python double loop.
//
// Here, per round: a CTA owns QB queries and one of 32 slices of the pixels, streams the feature map (coalesced channel
// planes, L2 resident: 13 MB, 8 loads in flight per thread), keeps (min distance, first index) and the maximum per query;
// a small kernel combines the slices into the round's winner = the next suppression window.  The previous winners'
// windows read as the map's maximum distance.  No [n,H,W] map is materialised.
// Distances are float32, channels accumulated in order without FMA contraction (what the oracle and torch's CPU
// reduction do).  argmax(exp(-d/2)) is resolved as argmin(d), first index on ties: identical unless two candidates'
// distances differ by less than the resolution of float32 exp (|dd| < 1.2e-7), documented in DESIGN.md.
#include <cuda_runtime.h>
#include <float.h>
#include <stdint.h>

#include "../../include/rp_b200.h"

namespace scnet { extern long long g_conv_launches; }

namespace kp {

constexpr int QB = 4, THREADS = 256, MAXC = 128, MAXK = 8, SPLITS = 32;

// workspace per query: windows [MAXK][2] int, dmax float, partials [SPLITS] x {d, idx, max}
struct Work { int* win; float* dmax; float* part_d; int* part_i; float* part_m; };
__host__ __device__ inline size_t work_bytes(int nq) { return (size_t)nq * (MAXK * 2 * 4 + 4 + SPLITS * 12) + 64; }
inline Work carve(void* ws, int nq) {
    Work w; char* p = static_cast<char*>(ws);
    w.win = reinterpret_cast<int*>(p); p += (size_t)nq * MAXK * 2 * 4;
    w.dmax = reinterpret_cast<float*>(p); p += (size_t)nq * 4;
    w.part_d = reinterpret_cast<float*>(p); p += (size_t)nq * SPLITS * 4;
    w.part_i = reinterpret_cast<int*>(p); p += (size_t)nq * SPLITS * 4;
    w.part_m = reinterpret_cast<float*>(p);
    return w;
}

// One round over one slice of the pixels for QB queries: (min distance, first index) and the maximum of the slice.
template <bool FROM_DIST>
__global__ void __launch_bounds__(THREADS) match_round_kernel(const float* __restrict__ q, int C, int nq, const float* __restrict__ feat,
                                                              const float* __restrict__ dist, int H, int W, int round, int window, Work wk) {
    __shared__ float sq[QB][MAXC];
    __shared__ float red_d[QB][THREADS / 32]; __shared__ int red_i[QB][THREADS / 32]; __shared__ float red_m[QB][THREADS / 32];
    __shared__ int win_x[QB][MAXK], win_y[QB][MAXK];
    __shared__ float s_dmax[QB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q0 = blockIdx.x * QB, split = blockIdx.y;
    const int HW = H * W;
    const int per = (HW + SPLITS - 1) / SPLITS, p_lo = split * per, p_hi = min(HW, p_lo + per);
    if (!FROM_DIST)
        for (int i = tid; i < QB * C; i += THREADS) { const int b = i / C, c = i - b * C; sq[b][c] = q0 + b < nq ? q[(size_t)c * nq + q0 + b] : 0.f; }
    if (tid < QB * MAXK) {
        const int b = tid / MAXK, j = tid - b * MAXK;
        if (q0 + b < nq && j < round) { win_x[b][j] = wk.win[((size_t)(q0 + b) * MAXK + j) * 2]; win_y[b][j] = wk.win[((size_t)(q0 + b) * MAXK + j) * 2 + 1]; }
    }
    if (tid < QB) s_dmax[tid] = (round > 0 && q0 + tid < nq) ? wk.dmax[q0 + tid] : 0.f;
    __syncthreads();
    float best[QB], dmx[QB]; int bidx[QB];
#pragma unroll
    for (int b = 0; b < QB; ++b) { best[b] = FLT_MAX; bidx[b] = 0x7fffffff; dmx[b] = -FLT_MAX; }
    for (int p = p_lo + tid; p < p_hi; p += THREADS) {
        float d[QB];
        if (FROM_DIST) {
#pragma unroll
            for (int b = 0; b < QB; ++b) d[b] = q0 + b < nq ? dist[(size_t)(q0 + b) * HW + p] : 0.f;
        } else {
#pragma unroll
            for (int b = 0; b < QB; ++b) d[b] = 0.f;
            int c = 0;
            for (; c + 8 <= C; c += 8) {                       // 8 channel loads in flight, accumulated in channel order
                float f[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) f[u] = __ldg(feat + (size_t)(c + u) * HW + p);
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int b = 0; b < QB; ++b) { const float t = __fsub_rn(sq[b][c + u], f[u]); d[b] = __fadd_rn(d[b], __fmul_rn(t, t)); }
            }
            for (; c < C; ++c) {
                const float f = __ldg(feat + (size_t)c * HW + p);
#pragma unroll
                for (int b = 0; b < QB; ++b) { const float t = __fsub_rn(sq[b][c], f); d[b] = __fadd_rn(d[b], __fmul_rn(t, t)); }
            }
        }
        const int y = p / W, x = p - y * W;
#pragma unroll
        for (int b = 0; b < QB; ++b) {
            float v = d[b];
            if (round == 0) dmx[b] = fmaxf(dmx[b], v);
            else {
                for (int j = 0; j < round; ++j) {               // suppressed windows take the map's maximum distance (= minimum heat)
                    const int wx = win_x[b][j], wy = win_y[b][j];
                    const int x_lo = max(0, wx - window), y_lo = max(0, wy - window);
                    const int x_hi = min(W - 1, wx + window), y_hi = min(H - 1, wy + window);      // exclusive (numpy slice)
                    if (x >= x_lo && x < x_hi && y >= y_lo && y < y_hi) v = s_dmax[b];
                }
            }
            if (v < best[b]) { best[b] = v; bidx[b] = p; }        // p increases per thread: strict < keeps the first index
        }
    }
#pragma unroll
    for (int b = 0; b < QB; ++b) {
        float bd = best[b]; int bi = bidx[b]; float bm = dmx[b];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const float om = __shfl_xor_sync(0xffffffffu, bm, o);
            if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            bm = fmaxf(bm, om);
        }
        if (lane == 0) { red_d[b][warp] = bd; red_i[b][warp] = bi; red_m[b][warp] = bm; }
    }
    __syncthreads();
    if (tid < QB && q0 + tid < nq) {
        const int b = tid;
        float bd = red_d[b][0]; int bi = red_i[b][0]; float bm = red_m[b][0];
        for (int w = 1; w < THREADS / 32; ++w) {
            if (red_d[b][w] < bd || (red_d[b][w] == bd && red_i[b][w] < bi)) { bd = red_d[b][w]; bi = red_i[b][w]; }
            bm = fmaxf(bm, red_m[b][w]);
        }
        const size_t o = (size_t)(q0 + b) * SPLITS + split;
        wk.part_d[o] = bd; wk.part_i[o] = bi; wk.part_m[o] = bm;
    }
}

// Combine the slices of one round: the winner becomes point `round` of the query and the next suppression window.
__global__ void match_finish_kernel(int nq, int W, int K, int round, Work wk, double* __restrict__ pts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    float bd = FLT_MAX; int bi = 0x7fffffff; float bm = -FLT_MAX;
    for (int s = 0; s < SPLITS; ++s) {
        const float d = wk.part_d[(size_t)i * SPLITS + s]; const int ix = wk.part_i[(size_t)i * SPLITS + s];
        if (d < bd || (d == bd && ix < bi)) { bd = d; bi = ix; }
        bm = fmaxf(bm, wk.part_m[(size_t)i * SPLITS + s]);
    }
    if (round == 0) wk.dmax[i] = bm;
    const int y = bi / W, x = bi - y * W;
    wk.win[((size_t)i * MAXK + round) * 2] = x; wk.win[((size_t)i * MAXK + round) * 2 + 1] = y;
    pts[((size_t)i * K + round) * 2] = (double)x; pts[((size_t)i * K + round) * 2 + 1] = (double)y;
}

template <bool FROM_DIST>
static int run(const float* q, int C, int nq, const float* feat, const float* dist, int H, int W, int K, int window, double* pts,
               void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (ws_bytes < work_bytes(nq)) return RP_ERR_WORKSPACE_TOO_SMALL;
    const Work wk = carve(ws, nq);
    dim3 grid((nq + QB - 1) / QB, SPLITS);
    for (int round = 0; round < K; ++round) {
        match_round_kernel<FROM_DIST><<<grid, THREADS, 0, stream>>>(q, C, nq, feat, dist, H, W, round, window, wk);
        match_finish_kernel<<<(nq + 127) / 128, 128, 0, stream>>>(nq, W, K, round, wk, pts);
        scnet::g_conv_launches += 2;
    }
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

}  // namespace kp

extern "C" {

int rp_match_sample_workspace_bytes(int nq, size_t* bytes) {
    if (nq < 0 || !bytes) return RP_ERR_INVALID_ARG;
    *bytes = kp::work_bytes(nq);
    return RP_OK;
}

int rp_match_sample(const float* q, int C, int nq, const float* feat, int H, int W, int K, int window, double* pts,
                    void* workspace, size_t workspace_bytes, void* stream_) {
    if (nq == 0) return RP_OK;
    if (!q || !feat || !pts || !workspace || C < 1 || C > kp::MAXC || nq < 0 || H < 1 || W < 1 || K < 1 || K > kp::MAXK || window < 0) return RP_ERR_INVALID_ARG;
    return kp::run<false>(q, C, nq, feat, nullptr, H, W, K, window, pts, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream_));
}

int rp_heat_sample(const float* dist, int n, int H, int W, int K, int window, double* pts, void* workspace, size_t workspace_bytes, void* stream_) {
    if (n == 0) return RP_OK;
    if (!dist || !pts || !workspace || n < 0 || H < 1 || W < 1 || K < 1 || K > kp::MAXK || window < 0) return RP_ERR_INVALID_ARG;
    return kp::run<true>(nullptr, 1, n, nullptr, dist, H, W, K, window, pts, workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream_));
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// Net -> solver hand-off on the device (rpmodule.getMatchingPrimitive, rpmodule.py:511-538, for fixed keypoint counts):
// per (image, keypoint) rputil.getPixel (rputil.py:61-119: bilinear depth / normal at the sub-pixel location, pinhole
// back-projection on the 160x640 four-face skybox, face rotation; float64) and rputil.interpolate of the descriptor map
// (rputil.py:43-58; float32, the reference's operation order), written as the rows [K,3] / [K,3] / [K,C] the solver reads.
namespace prim {

__global__ void gather_primitives_kernel(const float* __restrict__ feat, int C, long long feat_img_stride,
                                         const double* __restrict__ depth, const double* __restrict__ normal,
                                         const double* __restrict__ pts, int n_img, int K, int dataset,
                                         double* __restrict__ pc_out, double* __restrict__ nn_out, float* __restrict__ desc_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_img * K) return;
    const int im = idx / K;
    constexpr int H = 160, W = 640;
    const double px = pts[2 * (size_t)idx], py = pts[2 * (size_t)idx + 1];
    // ---- getPixel
    {
        const int tx = (int)floor(px), ty = (int)floor(py);
        const double fx = px - (double)tx, fy = py - (double)ty;
        const double w00 = (1.0 - fy) * (1.0 - fx), w01 = fx * (1.0 - fy), w10 = fy * (1.0 - fx), w11 = fx * fy;
        const double* d = depth + (size_t)im * H * W;
        const size_t p00 = (size_t)ty * W + tx, p01 = p00 + 1, p10 = p00 + W, p11 = p10 + 1;
        const double val = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(d[p00], w00), __dmul_rn(d[p01], w01)), __dmul_rn(d[p10], w10)), __dmul_rn(d[p11], w11));
        const double* nm = normal + (size_t)im * H * W * 3;
        double n[3];
#pragma unroll
        for (int c = 0; c < 3; ++c)
            n[c] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(nm[p00 * 3 + c], w00), __dmul_rn(nm[p01 * 3 + c], w01)), __dmul_rn(nm[p10 * 3 + c], w10)), __dmul_rn(nm[p11 * 3 + c], w11));
        const double nrm = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(n[0], n[0]), __dmul_rn(n[1], n[1])), __dmul_rn(n[2], n[2])));
#pragma unroll
        for (int c = 0; c < 3; ++c) nn_out[(size_t)idx * 3 + c] = n[c] / nrm;
        const int face = (int)floor(px / 160.0);
        const int r = dataset == 0 ? face : ((face + 3) & 3);
        const double y = __dmul_rn(__dmul_rn(__dsub_rn(0.5, py / 160.0), 2.0), val);
        const double x = __dmul_rn(__dmul_rn(__dsub_rn(__dsub_rn(px, (double)(face * 160)) / 160.0, 0.5), 2.0), val);
        const double z = -val;
        double ox, oz;
        if (r == 0) { ox = x; oz = z; } else if (r == 1) { ox = -z; oz = x; } else if (r == 2) { ox = -x; oz = -z; } else { ox = z; oz = -x; }
        pc_out[(size_t)idx * 3] = ox; pc_out[(size_t)idx * 3 + 1] = y; pc_out[(size_t)idx * 3 + 2] = oz;
    }
    // ---- interpolate (normalised coordinates pass through float32, as torch_op.v(ptsNorm) does)
    {
        const float xn = (float)(px / 640.0), yn = (float)(py / 160.0);
        const float x = xn * (float)(W - 1), y = yn * (float)(H - 1);
        const float x0 = floorf(x), y0 = floorf(y);
        int ix = (int)x0, iy = (int)y0;
        ix = ix < 0 ? 0 : (ix > W - 2 ? W - 2 : ix);
        iy = iy < 0 ? 0 : (iy > H - 2 ? H - 2 : iy);
        const float wx0 = __fsub_rn(__fadd_rn(x0, 1.f), x), wy0 = __fsub_rn(__fadd_rn(y0, 1.f), y);
        const float wx1 = __fsub_rn(x, x0), wy1 = __fsub_rn(y, y0);
        const float* fb = feat + (size_t)im * feat_img_stride;
        for (int c = 0; c < C; ++c) {
            const float* f = fb + (size_t)c * H * W;
            const float v00 = f[(size_t)iy * W + ix], v10 = f[(size_t)(iy + 1) * W + ix];
            const float v01 = f[(size_t)iy * W + ix + 1], v11 = f[(size_t)(iy + 1) * W + ix + 1];
            float rr = __fmul_rn(__fmul_rn(v00, wx0), wy0);
            rr = __fadd_rn(rr, __fmul_rn(__fmul_rn(v10, wx0), wy1));
            rr = __fadd_rn(rr, __fmul_rn(__fmul_rn(v01, wx1), wy0));
            rr = __fadd_rn(rr, __fmul_rn(__fmul_rn(v11, wx1), wy1));
            desc_out[(size_t)idx * C + c] = rr;
        }
    }
}

}  // namespace prim

extern "C" int rp_gather_primitives(const float* feat, int C, long long feat_img_stride, const double* depth, const double* normal,
                                    const double* pts, int n_img, int K, int dataset, double* pc_out, double* nn_out,
                                    float* desc_out, void* stream_) {
    if (n_img == 0 || K == 0) return RP_OK;
    if (!feat || !depth || !normal || !pts || !pc_out || !nn_out || !desc_out || C < 1 || n_img < 0 || K < 0 || dataset < 0 || dataset > 2)
        return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const int n = n_img * K;
    prim::gather_primitives_kernel<<<(n + 127) / 128, 128, 0, stream>>>(feat, C, feat_img_stride, depth, normal, pts, n_img, K, dataset,
                                                                        pc_out, nn_out, desc_out);
    ++scnet::g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

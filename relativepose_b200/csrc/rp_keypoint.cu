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

// scnet_tc.cu -- tcgen05 (5th-gen tensor core) building blocks for the SCNet convolutions, sm_100a only.
//
// Operands are staged in shared memory by ordinary threads (the A operand of an implicit-GEMM convolution is a
// gather with the producer's BatchNorm + LeakyReLU applied on the fly, so it cannot come from TMA) in the
// canonical K-major, no-swizzle UMMA layout: 8-row x 16-byte "core matrices", 128 contiguous bytes each;
//   core(kc, mc) of a [rows][BK] bf16 tile lives at ((kc * rows/8) + mc) * 128 bytes
//   -> stride between cores along M/N (SBO) = 128 B, along K (LBO) = rows/8 * 128 B.
// One elected thread issues tcgen05.mma (M=128, N=BN, K=16 per instruction, bf16 x bf16 -> fp32 in TMEM),
// tcgen05.commit signals an mbarrier, the epilogue reads the accumulator with tcgen05.ld (32 lanes x 32 columns
// per warp and instruction).  Descriptor bit layouts follow cute/arch/mma_sm100_desc.hpp (UMMA::SmemDescriptor,
// UMMA::InstrDescriptor).
#include "rp_h16.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rp_b200.h"
#include "scnet_common.cuh"

#include "tc_prims.cuh"

namespace tc {

constexpr int TM = 128;     // UMMA M
constexpr int TK = 64;      // K per pipeline stage (4 MMAs)

// ---------------------------------------------------------------------------------------------------
// Bring-up / unit-test kernel: C[M,N] = A[M,K] * B[N,K]^T, fp32 in/out, bf16 operands, fp32 accumulate.
// One CTA (128 threads) per 128 x BN tile; synchronous single-stage K loop.  M % 128 == 0, N % BN == 0, K % 64 == 0.
template <int BN>
__global__ void __launch_bounds__(128) gemm_bf16_test(const float* __restrict__ A, const float* __restrict__ B,
                                                      float* __restrict__ C, int M, int N, int K) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sA = smem;                       // 128 x 64 bf16 = 16 KB
    unsigned char* sB = smem + TM * TK * 2;         // BN  x 64 bf16
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * TM, n0 = blockIdx.y * BN;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_slot, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = make_idesc_bf16(TM, BN);
    uint32_t parity = 0;
    for (int k0 = 0; k0 < K; k0 += TK) {
        // stage A: thread = row
        {
            const float* ap = A + (size_t)(m0 + tid) * K + k0;
#pragma unroll
            for (int kc = 0; kc < TK / 8; ++kc) {
                float4 x = *reinterpret_cast<const float4*>(ap + kc * 8);
                float4 y = *reinterpret_cast<const float4*>(ap + kc * 8 + 4);
                float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                store_core_row(sA, TM, tid, kc, v);
            }
        }
        for (int r = tid; r < BN; r += 128) {
            const float* bp = B + (size_t)(n0 + r) * K + k0;
#pragma unroll
            for (int kc = 0; kc < TK / 8; ++kc) {
                float4 x = *reinterpret_cast<const float4*>(bp + kc * 8);
                float4 y = *reinterpret_cast<const float4*>(bp + kc * 8 + 4);
                float v[8] = {x.x, x.y, x.z, x.w, y.x, y.y, y.z, y.w};
                store_core_row(sB, BN, r, kc, v);
            }
        }
        fence_async_smem();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
            for (int j = 0; j < TK / 16; ++j) {
                uint64_t ad = make_smem_desc(a0 + j * 2 * (TM / 8) * 128, (TM / 8) * 128, 128);
                uint64_t bd = make_smem_desc(b0 + j * 2 * (BN / 8) * 128, (BN / 8) * 128, 128);
                umma_bf16(tmem_d, ad, bd, idesc, (k0 > 0 || j > 0) ? 1u : 0u);
            }
            umma_commit(&bar);       // arrives when the MMAs above have finished reading smem / writing TMEM
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
    }
    tc_fence_after();
    // epilogue: warp w reads TMEM lanes 32w..32w+31 (= rows), 32 columns at a time
    for (int c0 = 0; c0 < BN; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        float* cp = C + (size_t)(m0 + tid) * N + n0 + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(cp + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, BN);
}

// ---------------------------------------------------------------------------------------------------
// Implicit-GEMM convolution / transposed convolution on tcgen05.  One CTA (256 threads) = 128 output pixels (UMMA M)
// x BN output channels of one (scan pair, sub-pixel class); K loop over (tap, source, TK-channel tile) with NS=3
// shared-memory stages:
//   * B (weights): pre-packed on the host in the UMMA smem image, fetched by ONE thread with cp.async.bulk
//     (1-D TMA, completion on an mbarrier via complete_tx), issued one iteration ahead;
//   * A (activations): gathered by all 256 threads (2 per pixel row, TK/2 channels each) with the producer's
//     BatchNorm + LeakyReLU applied on the fly, converted to bf16 and written as 16-byte core-matrix rows;
//   * one thread issues tcgen05.mma; tcgen05.commit -> mbarrier frees the stage.
// Epilogue: TMEM -> registers -> raw fp32 NHWC output (+bias/tanh for the 1x1 heads) and per-channel partial batch
// statistics through a padded smem transpose (fixed summation order).
constexpr int NS = 3;
constexpr int CTA = 256;

// A-tile geometry: K-core stride (LBO) padded by 16 bytes so that the 8 lanes that stage the 8 K-cores of one pixel
// row hit 8 different 16-byte bank groups (conflict-free STS.128) while reading one contiguous 256-byte run of global.
constexpr int A_LBO = (TM / 8) * 128 + 16;
constexpr int MAX_CIN = 1536;

template <int BN, int TK>
__global__ void __launch_bounds__(CTA, 2) conv_igemm_tc(const scnet::ConvArgs A, const unsigned char* __restrict__ Wp,
                                                         int nkt, int ntn) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int KC = TK / 8;                            // 16-byte core rows (8 channels) per pixel row and stage
    constexpr int A_BYTES = ((KC * A_LBO + 127) / 128) * 128, B_BYTES = BN * TK * 2, STAGE = A_BYTES + B_BYTES;
    constexpr int NI = (TM * KC) / CTA;                   // (pixel row, core) items per thread and stage
    constexpr int ROWS_PER_PASS = CTA / KC;
    __shared__ __align__(8) uint64_t empty_bar[NS];
    __shared__ __align__(8) uint64_t fullb_bar[NS];
    __shared__ uint32_t tmem_slot;
    __shared__ float red_s[8][32], red_q[8][32];
    __shared__ __align__(16) float s_sc[MAX_CIN], s_sh[MAX_CIN];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row = ((warp & 3) << 5) + lane;             // epilogue: pixel row of the tile == TMEM lane (warps 0-3)
    const int kc_l = tid % KC;                            // loader: which 8-channel core of the K tile
    const int prow0 = tid / KC;                           // loader: first pixel row (then += ROWS_PER_PASS)
    const int tile_m = blockIdx.x, tile_n = blockIdx.y;
    const int g = blockIdx.z / A.nclass, ci = blockIdx.z - g * A.nclass;
    const scnet::ConvClass& C = A.cls[ci];
    const int HW = C.Ha * C.Wb;
    const int Mc = A.gsz * HW;
    const int part_row = (g * A.nclass + ci) * A.tiles_m + tile_m;
    if (tile_m * TM >= Mc) {             // padded tile of a smaller class: zeros for the statistics
        if (A.psum && tid < BN) {
            int co = tile_n * BN + tid;
            if (co < A.Cout) { A.psum[(size_t)part_row * A.Cout + co] = 0.f; A.psq[(size_t)part_row * A.Cout + co] = 0.f; }
        }
        return;
    }
    // loader pixel coordinates (NI rows per thread)
    int l_img[NI], l_a[NI], l_b[NI]; bool l_val[NI];
#pragma unroll
    for (int j = 0; j < NI; ++j) {
        const int m = tile_m * TM + prow0 + j * ROWS_PER_PASS;
        l_val[j] = m < Mc; l_img[j] = 0; l_a[j] = 0; l_b[j] = 0;
        if (l_val[j]) { int im = m / HW; int rem = m - im * HW; l_img[j] = g * A.gsz + im; l_a[j] = rem / C.Wb; l_b[j] = rem - l_a[j] * C.Wb; }
    }
    // producer BN scale/shift of every input channel of this group, once
    {
        int cb = 0;
        for (int si = 0; si < A.nsrc; ++si) {
            const rp_conv_src& S = A.src[si];
            for (int c = tid; c < S.C; c += CTA) {
                s_sc[cb + c] = S.act ? S.scale[(size_t)g * S.sstride + S.s_off + c] : 1.f;
                s_sh[cb + c] = S.act ? S.shift[(size_t)g * S.sstride + S.s_off + c] : 0.f;
            }
            cb += S.C;
        }
    }
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) { mbar_init(&empty_bar[i], 1); mbar_init(&fullb_bar[i], 1); }
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(&tmem_slot, BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    const uint32_t idesc = make_idesc_bf16(TM, BN);
    const int nkt0 = A.src[0].C / TK;
    const int niter = C.ntap * nkt;

    auto weight_block = [&](int i) -> const unsigned char* {
        const int t = i / nkt, kt = i - t * nkt;
        return Wp + ((size_t)((size_t)C.taps[t].widx * nkt + kt) * ntn + tile_n) * B_BYTES;
    };
    if (tid == 0) {                                       // prologue: weights of iteration 0
        mbar_expect_tx(&fullb_bar[0], B_BYTES);
        bulk_g2s(smem + A_BYTES, weight_block(0), B_BYTES, &fullb_bar[0]);
    }

    // raw A registers for the current and the next iteration (software pipelining of the gather); 8 channels per
    // item = 32 bytes of float32 or 16 bytes of bfloat16 storage
    uint4 xc[NI][2], xn[NI][2];
    bool vc[NI], vn[NI];
    auto issue_loads = [&](int i, uint4 (&x)[NI][2], bool (&v)[NI]) {
        const int t = i / nkt, kt = i - t * nkt;
        const scnet::Tap tp = C.taps[t];
        const int si = kt < nkt0 ? 0 : 1;
        const rp_conv_src& S = A.src[si];
        const int c0 = (kt - (si ? nkt0 : 0)) * TK + kc_l * 8;
        const bool h16 = S.dtype == 1;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int iy = l_a[j] * A.istr + tp.dy, ix = l_b[j] * A.istr + tp.dx;
            v[j] = l_val[j] && iy >= 0 && iy < A.Hin && ix >= 0 && ix < A.Win;
            if (v[j]) {
                const size_t e = (((size_t)l_img[j] * A.Hin + iy) * A.Win + ix) * S.pitch + S.ch_off + c0;
                if (h16) {
                    x[j][0] = *reinterpret_cast<const uint4*>(reinterpret_cast<const rp_h16*>(S.ptr) + e);
                } else {
                    const uint4* p = reinterpret_cast<const uint4*>(S.ptr + e);
                    x[j][0] = p[0];
                    x[j][1] = p[1];
                }
            }
        }
    };
    issue_loads(0, xc, vc);

    for (int i = 0; i < niter; ++i) {
        const int stage = i % NS, use = i / NS;
        unsigned char* sA = smem + stage * STAGE;
        unsigned char* sB = sA + A_BYTES;
        if (tid == 0 && i + 1 < niter) {                  // prefetch the next iteration's weight block (bulk TMA)
            const int s1 = (i + 1) % NS, u1 = (i + 1) / NS;
            mbar_wait(&empty_bar[s1], (uint32_t)((u1 & 1) ^ 1));
            mbar_expect_tx(&fullb_bar[s1], B_BYTES);
            bulk_g2s(smem + s1 * STAGE + A_BYTES, weight_block(i + 1), B_BYTES, &fullb_bar[s1]);
        }
        if (i + 1 < niter) issue_loads(i + 1, xn, vn);    // next A gather in flight while this tile is transformed
        const int kt = i % nkt;
        const int cb = kt * TK + kc_l * 8;                // channel index into s_sc/s_sh (sources are concatenated)
        const float4 s0 = *reinterpret_cast<const float4*>(&s_sc[cb]), s1v = *reinterpret_cast<const float4*>(&s_sc[cb + 4]);
        const float4 h0 = *reinterpret_cast<const float4*>(&s_sh[cb]), h1 = *reinterpret_cast<const float4*>(&s_sh[cb + 4]);
        const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1v.x, s1v.y, s1v.z, s1v.w};
        const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const bool act = A.src[kt < nkt0 ? 0 : 1].act != 0;
        const float slope = A.src[kt < nkt0 ? 0 : 1].slope;
        const bool h16 = A.src[kt < nkt0 ? 0 : 1].dtype == 1;
        mbar_wait(&empty_bar[stage], (uint32_t)((use & 1) ^ 1));     // MMAs that read this stage are done
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int prow = prow0 + j * ROWS_PER_PASS;
            uint4 u = make_uint4(0u, 0u, 0u, 0u);
            if (vc[j]) {
                float v[8];
                if (h16) {
                    const rp_h162* hp = reinterpret_cast<const rp_h162*>(&xc[j][0]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) { const float2 f = rp_h2_to_f2(hp[q]); v[2 * q] = f.x; v[2 * q + 1] = f.y; }
                } else {
                    v[0] = __uint_as_float(xc[j][0].x); v[1] = __uint_as_float(xc[j][0].y); v[2] = __uint_as_float(xc[j][0].z);
                    v[3] = __uint_as_float(xc[j][0].w); v[4] = __uint_as_float(xc[j][1].x); v[5] = __uint_as_float(xc[j][1].y);
                    v[6] = __uint_as_float(xc[j][1].z); v[7] = __uint_as_float(xc[j][1].w);
                }
                if (act) {
#pragma unroll
                    for (int q = 0; q < 8; ++q) { float z = fmaf(v[q], sv[q], hv[q]); v[q] = z > 0.f ? z : slope * z; }
                }
                rp_h162 p0 = rp_f2_to_h2(v[0], v[1]), p1 = rp_f2_to_h2(v[2], v[3]);
                rp_h162 p2 = rp_f2_to_h2(v[4], v[5]), p3 = rp_f2_to_h2(v[6], v[7]);
                u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
                u.z = *reinterpret_cast<uint32_t*>(&p2); u.w = *reinterpret_cast<uint32_t*>(&p3);
            }
            *reinterpret_cast<uint4*>(sA + kc_l * A_LBO + (prow >> 3) * 128 + (prow & 7) * 16) = u;
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            mbar_wait(&fullb_bar[stage], (uint32_t)(use & 1));      // this stage's weight block has landed
            tc_fence_after();
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
#pragma unroll
            for (int j = 0; j < TK / 16; ++j) {
                uint64_t ad = make_smem_desc(a0 + j * 2 * A_LBO, A_LBO, 128);
                uint64_t bd = make_smem_desc(b0 + j * 2 * (BN / 8) * 128, (BN / 8) * 128, 128);
                umma_bf16(tmem_d, ad, bd, idesc, (i > 0 || j > 0) ? 1u : 0u);
            }
            umma_commit(&empty_bar[stage]);
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) { xc[j][0] = xn[j][0]; xc[j][1] = xn[j][1]; vc[j] = vn[j]; }
    }
    {   // drain: the last commit covers every MMA issued before it
        const int last = niter - 1;
        mbar_wait(&empty_bar[last % NS], (uint32_t)((last / NS) & 1));
    }
    tc_fence_after();
    __syncthreads();           // every thread is past the mainloop: stage memory can be reused

    // ---- epilogue (warps 0-3 own the TMEM lanes; all 8 warps help with the column sums)
    float* Tt = reinterpret_cast<float*>(smem);            // [128][33] transpose buffer
    float* op = nullptr;
    rp_h16* oph = nullptr;
    if (warp < 4) {
        const int m_l = tile_m * TM + row;
        if (m_l < Mc) {
            const int im = m_l / HW; const int rem = m_l - im * HW; const int a_l = rem / C.Wb, b_l = rem - a_l * C.Wb;
            const int oy = a_l * A.ostr + C.py, ox = b_l * A.ostr + C.px;
            const size_t e = (((size_t)(g * A.gsz + im) * A.Hout + oy) * A.Wout + ox) * A.out_pitch + A.out_ch_off;
            if (A.out_bf16) oph = reinterpret_cast<rp_h16*>(A.out) + e; else op = A.out + e;
        }
    }
    for (int c0 = 0; c0 < BN; c0 += 32) {
        const int co0 = tile_n * BN + c0;
        if (warp < 4) {
            float v[32];
            tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            if (A.bias || A.tanh_out) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (co0 + j < A.Cout) { float y = v[j] + (A.bias ? A.bias[co0 + j] : 0.f); v[j] = A.tanh_out ? tanhf(y) : y; }
                }
            }
            if (A.out_bf16) {       // round to the storage type first: the statistics describe what the consumer reads
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = rp_h_to_f(rp_f_to_h(v[j]));
                if (oph) {
                    if (co0 + 31 < A.Cout && (((size_t)(oph + co0)) & 15) == 0) {
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            rp_h162 p0 = rp_f2_to_h2(v[j], v[j + 1]), p1 = rp_f2_to_h2(v[j + 2], v[j + 3]);
                            rp_h162 p2 = rp_f2_to_h2(v[j + 4], v[j + 5]), p3 = rp_f2_to_h2(v[j + 6], v[j + 7]);
                            uint4 o;
                            o.x = *reinterpret_cast<uint32_t*>(&p0); o.y = *reinterpret_cast<uint32_t*>(&p1);
                            o.z = *reinterpret_cast<uint32_t*>(&p2); o.w = *reinterpret_cast<uint32_t*>(&p3);
                            *reinterpret_cast<uint4*>(oph + co0 + j) = o;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) if (co0 + j < A.Cout) oph[co0 + j] = rp_f_to_h(v[j]);
                    }
                }
            }
            if (op) {
                if (co0 + 31 < A.Cout && (((size_t)(op + co0)) & 15) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(op + co0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) if (co0 + j < A.Cout) op[co0 + j] = v[j];
                }
            }
            if (A.psum) {
#pragma unroll
                for (int j = 0; j < 32; ++j) Tt[row * 33 + j] = v[j];          // rows of invalid pixels are exact zeros
            }
        }
        if (A.psum) {
            __syncthreads();
            {
                const int col = tid & 31, part = tid >> 5;                      // 8 parts of 16 rows
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int r = 0; r < 16; ++r) { float x = Tt[(part * 16 + r) * 33 + col]; s1 += x; s2 += x * x; }
                red_s[part][col] = s1; red_q[part][col] = s2;
            }
            __syncthreads();
            if (tid < 32 && co0 + tid < A.Cout) {
                float s1 = 0.f, s2 = 0.f;
#pragma unroll
                for (int r = 0; r < 8; ++r) { s1 += red_s[r][tid]; s2 += red_q[r][tid]; }
                A.psum[(size_t)part_row * A.Cout + co0 + tid] = s1;
                A.psq[(size_t)part_row * A.Cout + co0 + tid] = s2;
            }
            __syncthreads();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, BN);
}

template <int BN, int TK>
int launch_conv_tc(const scnet::ConvArgs& A, const void* wp, int nkt, cudaStream_t stream) {
    const int ntn = (A.Cout + BN - 1) / BN;
    size_t pipe = (size_t)NS * (size_t)((((TK / 8) * A_LBO + 127) / 128) * 128 + BN * TK * 2), tr = (size_t)128 * 33 * 4;
    size_t smem = pipe > tr ? pipe : tr;
    auto kern = conv_igemm_tc<BN, TK>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
    dim3 grid(A.tiles_m, ntn, A.G * A.nclass);
    kern<<<grid, CTA, smem, stream>>>(A, static_cast<const unsigned char*>(wp), nkt, ntn);
    ++scnet::g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

}  // namespace tc

extern "C" {

// 16-bit operand / activation format the library was built with: 1 = IEEE half, 0 = bfloat16 (rp_h16.cuh).
int rp_h16_format(void) { return RP_H16_FP16 ? 1 : 0; }


// Unit-test hook for the tcgen05 building blocks: C = A * B^T with bf16-rounded operands (device pointers).
int rp_tc_gemm_test(const float* A, const float* B, float* C, int M, int N, int K, int bn, void* stream_) {
    if (!A || !B || !C || M % 128 || K % 64 || (bn != 64 && bn != 128) || N % bn) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    dim3 grid(M / 128, N / bn);
    size_t smem = (size_t)(128 + bn) * 64 * 2;
    if (bn == 64) tc::gemm_bf16_test<64><<<grid, 128, smem, stream>>>(A, B, C, M, N, K);
    else tc::gemm_bf16_test<128><<<grid, 128, smem, stream>>>(A, B, C, M, N, K);
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_conv_nparts_tc(const rp_conv_desc* d, int* nparts) {
    scnet::ConvArgs A;
    if (!nparts || !scnet::build_args(d, &A, tc::TM)) return RP_ERR_INVALID_ARG;
    *nparts = A.nclass * A.tiles_m;
    return RP_OK;
}

// Same layer contract as rp_conv_layer, on tcgen05: bf16 operands, fp32 accumulation in TMEM.  `w_packed` = the
// layer's weights as bf16 blocks [tap][k-tile][n-tile][tk/8][bn/8][8 rows (co)][8 (ci)] (the UMMA smem image, see
// relativepose_b200/scnet_engine.py:pack_tc); every source must have C % tk == 0; bn in {32,64,128}, tk in {32,64}.
int rp_conv_layer_tc(const rp_conv_desc* d, const void* w_packed, int bn, int tk, void* stream_) {
    scnet::ConvArgs A;
    if (!w_packed || !scnet::build_args(d, &A, tc::TM)) return RP_ERR_INVALID_ARG;
    if (!d->out || !d->src[0].ptr) return RP_ERR_INVALID_ARG;
    int nkt = 0;
    if (A.Cin_total > tc::MAX_CIN) return RP_ERR_UNSUPPORTED;
    for (int i = 0; i < d->nsrc; ++i) {
        const int al = d->src[i].dtype == 1 ? 8 : 4;
        if (d->src[i].C % tk || (d->src[i].pitch % al) || (d->src[i].ch_off % al)) return RP_ERR_UNSUPPORTED;
        if (d->src[i].act && ((d->src[i].sstride % 4) || (d->src[i].s_off % 4))) return RP_ERR_UNSUPPORTED;
        nkt += d->src[i].C / tk;
    }
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
#define RP_TC_CASE(BN_, TK_) if (bn == BN_ && tk == TK_) return tc::launch_conv_tc<BN_, TK_>(A, w_packed, nkt, stream);
    RP_TC_CASE(32, 32) RP_TC_CASE(64, 32) RP_TC_CASE(128, 32) RP_TC_CASE(32, 64) RP_TC_CASE(64, 64) RP_TC_CASE(128, 64)
#undef RP_TC_CASE
    return RP_ERR_UNSUPPORTED;
}

}  // extern "C"

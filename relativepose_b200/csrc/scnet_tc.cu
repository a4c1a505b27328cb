// scnet_tc.cu -- bring-up GEMM for the tcgen05 (5th-gen tensor core) building blocks of tc_prims.cuh, sm_100a only.
// (The convolutions themselves live in scnet_halo.cu; the per-tap implicit-GEMM kernel that used to live here was removed
// in round 2 once the halo kernel covered every tensor-core layer shape.)
//
// Operands are staged in shared memory by ordinary threads in the canonical K-major, no-swizzle UMMA layout: 8-row x 16-byte "core matrices", 128 contiguous bytes each;
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

}  // extern "C"

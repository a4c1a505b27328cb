// rp_host.cu -- host-buffer entry for ONE scan pair: the call pattern of the reference's evaluation loop
// (evaluation.py:278-284 -> RelativePoseEstimation_helper, RPModule/rpmodule.py:317-508, one pair per call).
//
// rp_solve_pair_host takes the caller's host arrays as they are, packs them into one page-locked staging block (256-byte
// aligned sections), issues ONE host-to-device copy, the fused solver launch (rp_solve_batch_ex, B = 1, one workspace slot) and
// ONE device-to-host copy of pose + status + stats, and waits on the stream.  Everything a call needs on the device -- staging
// block, workspace, parameter block -- is cached per device and only grows, so the steady-state cost above the kernel is two
// small DMA transfers and a stream synchronise (the Python layer above it only passes pointers).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

#include "../../include/rp_b200.h"

namespace {

struct HostCtx {
    unsigned char* pin = nullptr;      // page-locked staging: inputs, then the 256-byte result block
    unsigned char* dev = nullptr;
    size_t cap = 0;
    void* ws = nullptr;
    size_t ws_bytes = 0;
};

HostCtx g_ctx[64];
std::mutex g_mu;

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

extern "C" int rp_solve_pair_host(int ns, int nt,
                                  const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                                  const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                                  int feat_dim, const rp_params* params, const int32_t* zero_row_topk, int max_topk,
                                  int feat_sum_order, int64_t edge_cap,
                                  double* T_out, int32_t* status, int32_t* stats, void* stream_) {
    if (ns < 1 || nt < 1 || !pc_s || !nrm_s || !feat_s || !w_s || !pc_t || !nrm_t || !feat_t || !w_t || !params || !T_out || !status)
        return RP_ERR_INVALID_ARG;
    if (max_topk < 1 || max_topk > RP_MAX_TOPK || feat_dim < 1 || feat_dim > RP_MAX_FEAT_DIM) return RP_ERR_UNSUPPORTED;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return RP_ERR_NO_DEVICE; }
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    std::lock_guard<std::mutex> lock(g_mu);
    HostCtx& C = g_ctx[dev];

    // staging layout (every section 256-byte aligned)
    const size_t b_hdr = up256(sizeof(int32_t) * (5 + (size_t)max_topk));      // off_s[2] off_t[2] order[1] zero_row_topk[max_topk]
    const size_t b_par = up256(sizeof(rp_params));
    const size_t b_p3s = up256(sizeof(double) * 3 * (size_t)ns), b_p3t = up256(sizeof(double) * 3 * (size_t)nt);
    const size_t b_ws = up256(sizeof(double) * (size_t)ns), b_wt = up256(sizeof(double) * (size_t)nt);
    const size_t b_fs = up256(sizeof(float) * (size_t)feat_dim * ns), b_ft = up256(sizeof(float) * (size_t)feat_dim * nt);
    const size_t o_hdr = 0, o_par = o_hdr + b_hdr, o_pcs = o_par + b_par, o_nrs = o_pcs + b_p3s, o_wss = o_nrs + b_p3s;
    const size_t o_pct = o_wss + b_ws, o_nrt = o_pct + b_p3t, o_wst = o_nrt + b_p3t, o_fs = o_wst + b_wt, o_ft = o_fs + b_fs;
    const size_t in_bytes = o_ft + b_ft, o_out = in_bytes, total = in_bytes + 256;     // result: T (128 B) status (4) stats (32)
    if (total > C.cap) {
        if (cudaStreamSynchronize(stream) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
        if (C.pin) cudaFreeHost(C.pin);
        if (C.dev) cudaFree(C.dev);
        C.pin = nullptr; C.dev = nullptr; C.cap = 0;
        const size_t cap = total * 2 > (1u << 16) ? total * 2 : (1u << 16);
        if (cudaMallocHost(reinterpret_cast<void**>(&C.pin), cap) != cudaSuccess || cudaMalloc(reinterpret_cast<void**>(&C.dev), cap) != cudaSuccess) {
            cudaGetLastError();
            if (C.pin) { cudaFreeHost(C.pin); C.pin = nullptr; }
            return RP_ERR_CUDA;
        }
        C.cap = cap;
    }
    size_t need_ws = 0;
    int rc = rp_solve_workspace_bytes(1, ns, nt, max_topk, feat_dim, edge_cap, &need_ws);
    if (rc != RP_OK) return rc;
    if (need_ws > C.ws_bytes) {
        if (cudaStreamSynchronize(stream) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
        if (C.ws) cudaFree(C.ws);
        C.ws = nullptr; C.ws_bytes = 0;
        const size_t cap = need_ws + need_ws / 4;
        if (cudaMalloc(&C.ws, cap) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
        C.ws_bytes = cap;
    }

    int32_t* hdr = reinterpret_cast<int32_t*>(C.pin + o_hdr);
    hdr[0] = 0; hdr[1] = ns; hdr[2] = 0; hdr[3] = nt; hdr[4] = feat_sum_order ? 1 : 0;
    for (int k = 0; k < max_topk; ++k) hdr[5 + k] = zero_row_topk ? zero_row_topk[k] : -1;
    memcpy(C.pin + o_par, params, sizeof(rp_params));
    memcpy(C.pin + o_pcs, pc_s, sizeof(double) * 3 * (size_t)ns);
    memcpy(C.pin + o_nrs, nrm_s, sizeof(double) * 3 * (size_t)ns);
    memcpy(C.pin + o_wss, w_s, sizeof(double) * (size_t)ns);
    memcpy(C.pin + o_pct, pc_t, sizeof(double) * 3 * (size_t)nt);
    memcpy(C.pin + o_nrt, nrm_t, sizeof(double) * 3 * (size_t)nt);
    memcpy(C.pin + o_wst, w_t, sizeof(double) * (size_t)nt);
    memcpy(C.pin + o_fs, feat_s, sizeof(float) * (size_t)feat_dim * ns);
    memcpy(C.pin + o_ft, feat_t, sizeof(float) * (size_t)feat_dim * nt);
    if (cudaMemcpyAsync(C.dev, C.pin, in_bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }

    unsigned char* d = C.dev;
    const int32_t* d_hdr = reinterpret_cast<const int32_t*>(d + o_hdr);
    double* dT = reinterpret_cast<double*>(d + o_out);
    int32_t* dstatus = reinterpret_cast<int32_t*>(d + o_out + 128);
    int32_t* dstats = reinterpret_cast<int32_t*>(d + o_out + 160);
    rc = rp_solve_batch_ex(1, d_hdr, d_hdr + 2,
                           reinterpret_cast<const double*>(d + o_pcs), reinterpret_cast<const double*>(d + o_nrs),
                           reinterpret_cast<const float*>(d + o_fs), reinterpret_cast<const double*>(d + o_wss),
                           reinterpret_cast<const double*>(d + o_pct), reinterpret_cast<const double*>(d + o_nrt),
                           reinterpret_cast<const float*>(d + o_ft), reinterpret_cast<const double*>(d + o_wst),
                           feat_dim, reinterpret_cast<const rp_params*>(d + o_par), nullptr, d_hdr + 5, d_hdr + 4,
                           ns, nt, max_topk, 1, edge_cap, C.ws, C.ws_bytes, dT, dstatus, dstats, RP_STAGE_SOLVE, nullptr, stream);
    if (rc != RP_OK) return rc;
    if (cudaMemcpyAsync(C.pin + o_out, d + o_out, 256, cudaMemcpyDeviceToHost, stream) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
    if (cudaStreamSynchronize(stream) != cudaSuccess) { cudaGetLastError(); return RP_ERR_CUDA; }
    memcpy(T_out, C.pin + o_out, 128);
    memcpy(status, C.pin + o_out + 128, sizeof(int32_t));
    if (stats) memcpy(stats, C.pin + o_out + 160, sizeof(int32_t) * RP_STATS_STRIDE);
    return RP_OK;
}

// rp_warp.cu -- view warping between the two scans of a pair, sm_100a.
//
// Reference: util.warping (util.py:94-172) = depth2pc / Pano2PointCloud (util.py:468-523, 751-811) -> rigid transform ->
// reproj_helper (util.py:537-749) called three times (colour, normal, depth), each a numpy fancy-index scatter
// `proj[y, x] = value` whose duplicates resolve "last write wins".  The observed pixels of a view are lifted to 3-D,
// moved by R and splatted onto the four skybox faces of the other view.
//
// Here: one scatter kernel and one gather kernel for a whole batch of views.
//   * A point hits at most one face (|x/z| < 1, |y/z| < 1, z < 0 are exclusive between the four face frames) and the
//     four faces own disjoint column blocks, so the reference's face order (front, left, back, right) never decides
//     anything; only "the later source pixel wins" does.  Source order = raster order of the observed window, so the
//     winner of a target pixel is max(source raster index): one atomicMax per source pixel into an int32 map.
//   * The gather kernel re-derives the winner's values from the source view (same deterministic arithmetic), so no
//     payload is scattered and the result does not depend on thread scheduling.
// Arithmetic is float64 in numpy's order; the 4x4 / 3x3 matmuls are FMA chains over k (what numpy's dgemm does on
// the build host: verified bit for bit in tests/golden/make_warp_golden.py's generator notes).  HBM-bound and tiny:
// 33 B read + 32 B written per target pixel.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/rp_b200.h"

namespace warp {

constexpr int H = 160, W = 640, HW = H * W;

struct Window { int x0, y0, w, h; };
__host__ __device__ inline Window window_of(int ds) {
    Window wd;
    if (ds == 2) { wd.x0 = 160 + 80 - 44; wd.y0 = 80 - 33; wd.w = 88; wd.h = 66; }     // util.py:148 (Kinect window)
    else { wd.x0 = 160; wd.y0 = 0; wd.w = 160; wd.h = 160; }                               // util.py:117-120, 132-136
    return wd;
}

__device__ __forceinline__ bool is_identity(const double* R) {
#pragma unroll
    for (int i = 0; i < 16; ++i) if (R[i] != ((i % 5 == 0) ? 1.0 : 0.0)) return false;
    return true;
}

// 3-D point of source pixel (lx, ly) of the observed window, already moved into the target frame.  Returns false for
// pixels the reference drops (depth == 0; not for suncg, whose Pano2PointCloud "assumes depth clean", util.py:768).
__device__ __forceinline__ bool target_point(int ds, const float* __restrict__ vb, const double* __restrict__ R, int lx, int ly,
                                             const Window wd, double P[3]) {
    const float zf = vb[6 * HW + (wd.y0 + ly) * W + wd.x0 + lx];
    if (ds != 0 && zf == 0.f) return false;
    const double z = (double)zf;
    double px, py, pz;
    if (ds == 2) {             // depth2pc, 66x88 branch (util.py:509-519)
        const double xs = __dmul_rn(__dsub_rn(__ddiv_rn((double)lx, 88.0), 0.5), 2.0);
        const double ys = __dmul_rn(__dsub_rn(0.5, __ddiv_rn((double)ly, 66.0)), 2.0);
        px = __ddiv_rn(__dmul_rn(__dmul_rn(xs, z), 88.0), 160.0);
        py = __ddiv_rn(__dmul_rn(__dmul_rn(ys, z), 66.0), 160.0);
        pz = -z;
    } else {
        const double xs = __dmul_rn(__dsub_rn(__ddiv_rn((double)lx, 160.0), 0.5), 2.0);
        const double ys = __dmul_rn(__dsub_rn(0.5, __ddiv_rn((double)ly, 160.0)), 2.0);
        const double vx = __dmul_rn(xs, z), vy = __dmul_rn(ys, z), vz = -z;
        if (ds == 0) { px = -vz; py = vy; pz = vx; }      // Rs[1] @ v: the observed face is the second one (util.py:483-484, 756-772)
        else { px = vx; py = vy; pz = vz; }               // matterport: depth2pc leaves the face frame (util.py:485-498)
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
        P[i] = fma(R[4 * i + 3], 1.0, fma(R[4 * i + 2], pz, fma(R[4 * i + 1], py, __dmul_rn(R[4 * i], px))));
    return true;
}

// tp = Rs[face]^T @ P for the panorama column block `slot` (util.py:546-602 suncg, :615-672 matterport, :686-743 scannet)
__device__ __forceinline__ void face_frame(int ds, int slot, const double P[3], double& t0, double& t1, double& t2) {
    const int face = ds == 0 ? slot : ((slot + 3) & 3);
    t1 = P[1];
    if (face == 0) { t0 = P[0]; t2 = P[2]; }
    else if (face == 1) { t0 = P[2]; t2 = -P[0]; }
    else if (face == 2) { t0 = -P[0]; t2 = -P[2]; }
    else { t0 = -P[2]; t2 = P[0]; }
}

__device__ __forceinline__ bool project(int ds, const double P[3], int& ox, int& oy, int& slot_out, double& depth) {
#pragma unroll
    for (int slot = 0; slot < 4; ++slot) {
        double t0, t1, t2;
        face_frame(ds, slot, P, t0, t1, t2);
        const double den = __dadd_rn(fabs(t2), 1e-32);
        const double u = __ddiv_rn(t0, den), v = __ddiv_rn(t1, den);
        if (t2 < 0.0 && fabs(u) < 1.0 && fabs(v) < 1.0) {
            double cx = rint(__dmul_rn(__dmul_rn(__dadd_rn(u, 1.0), 0.5), 160.0));      // np.round = half to even
            double cy = rint(__dmul_rn(__dmul_rn(__dsub_rn(1.0, v), 0.5), 160.0));
            cx = fmin(fmax(cx, 0.0), 159.0); cy = fmin(fmax(cy, 0.0), 159.0);
            ox = (int)cx + 160 * slot; oy = (int)cy; slot_out = slot; depth = -t2;
            return true;
        }
    }
    return false;
}

__global__ void warp_scatter_kernel(const float* __restrict__ view, long long view_stride, const int* __restrict__ src_index,
                                    const double* __restrict__ Rall, int B, int ds, int* __restrict__ zbuf) {
    const Window wd = window_of(ds);
    const int npx = wd.w * wd.h;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * npx) return;
    const int b = idx / npx, j = idx - b * npx;
    const double* R = Rall + 16 * b;
    if (is_identity(R)) return;
    const int ly = j / wd.w, lx = j - ly * wd.w;
    double P[3];
    if (!target_point(ds, view + (size_t)(src_index ? src_index[b] : b) * view_stride, R, lx, ly, wd, P)) return;
    int ox, oy, slot; double depth;
    if (!project(ds, P, ox, oy, slot, depth)) return;
    atomicMax(&zbuf[(size_t)b * HW + oy * W + ox], j);
}

__global__ void warp_gather_kernel(const float* __restrict__ view, long long view_stride, const int* __restrict__ src_index,
                                   const double* __restrict__ Rall, int B, int ds, const int* __restrict__ zbuf,
                                   float* __restrict__ out, long long out_stride) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * HW) return;
    const int b = idx / HW, t = idx - b * HW;
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const double* R = Rall + 16 * b;
    const int j = is_identity(R) ? -1 : zbuf[idx];
    if (j >= 0) {
        const Window wd = window_of(ds);
        const float* vb = view + (size_t)(src_index ? src_index[b] : b) * view_stride;
        const int ly = j / wd.w, lx = j - ly * wd.w;
        const int sp = (wd.y0 + ly) * W + wd.x0 + lx;
        double P[3];
        target_point(ds, vb, R, lx, ly, wd, P);
        int ox, oy, slot; double depth;
        project(ds, P, ox, oy, slot, depth);
        const double nx = (double)vb[3 * HW + sp], ny = (double)vb[4 * HW + sp], nz = (double)vb[5 * HW + sp];
        o[0] = vb[sp]; o[1] = vb[HW + sp]; o[2] = vb[2 * HW + sp];
#pragma unroll
        for (int i = 0; i < 3; ++i) o[3 + i] = (float)fma(R[4 * i + 2], nz, fma(R[4 * i + 1], ny, __dmul_rn(R[4 * i], nx)));
        o[6] = (float)depth;
        o[7] = depth != 0.0 ? 1.f : 0.f;
    }
    float* ob = out + (size_t)b * out_stride + t;
#pragma unroll
    for (int c = 0; c < 8; ++c) ob[(size_t)c * HW] = o[c];
}

// util.Pano2PointCloud (util.py:751-811), dense: pc [B,3,102400] float64 in the reference's order (face, row, column);
// valid[B,102400] = 1 where the reference keeps the point (scannet drops depth == 0; the others keep everything).
__global__ void pano2pc_kernel(const float* __restrict__ depth, int B, int ds, double* __restrict__ pc, unsigned char* __restrict__ valid) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * HW) return;
    const int b = idx / HW, k = idx - b * HW;
    const int face = k / (160 * 160), r = k - face * 160 * 160, y = r / 160, x = r - y * 160;
    const float zf = depth[(size_t)b * HW + y * W + face * 160 + x];
    const double z = (double)zf;
    const double xs = __dmul_rn(__dsub_rn(__ddiv_rn((double)x, 160.0), 0.5), 2.0);
    const double ys = __dmul_rn(__dsub_rn(0.5, __ddiv_rn((double)y, 160.0)), 2.0);
    double vx, vy;
    if (ds == 2) { vx = __ddiv_rn(__dmul_rn(xs, z), 0.8921875 * 2); vy = __ddiv_rn(__dmul_rn(ys, z), 1.1895 * 2); }
    else { vx = __dmul_rn(xs, z); vy = __dmul_rn(ys, z); }
    const double vz = -z;
    const int rot = ds == 0 ? face : ((face + 3) & 3);
    double px, pz;
    if (rot == 0) { px = vx; pz = vz; }
    else if (rot == 1) { px = -vz; pz = vx; }
    else if (rot == 2) { px = -vx; pz = -vz; }
    else { px = vz; pz = -vx; }
    const bool keep = ds != 2 || zf != 0.f;
    double* o = pc + (size_t)b * 3 * HW + k;
    o[0] = keep ? px : 0.0; o[HW] = keep ? vy : 0.0; o[2 * (size_t)HW] = keep ? pz : 0.0;
    if (valid) valid[idx] = keep ? 1 : 0;
}

// Blend of RelativePoseEstimationViaCompletion (rpmodule.py:628-634): observed region from the input scan, the rest from
// the network; normals re-normalised with EPS = 1e-12.  T = dtype of the caller's normal / depth arrays (numpy promotes
// float32 * T -> T).
template <typename T>
__global__ void blend_kernel(const float* __restrict__ f, int C, const float* __restrict__ mask, const T* __restrict__ norm_gt,
                             const T* __restrict__ depth_gt, int B, T* __restrict__ normal_out, T* __restrict__ depth_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * HW) return;
    const int b = idx / HW, t = idx - b * HW;
    const float m = mask[idx];
    const float om = 1.f - m;
    const float* fb = f + (size_t)b * C * HW + t;
    T s[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) s[c] = (T)(om * fb[(size_t)(3 + c) * HW]) + (T)m * norm_gt[(size_t)idx * 3 + c];
    const T nn = sqrt((s[0] * s[0] + s[1] * s[1]) + s[2] * s[2]) + (T)1e-12;
#pragma unroll
    for (int c = 0; c < 3; ++c) normal_out[(size_t)idx * 3 + c] = s[c] / nn;
    depth_out[idx] = (T)(om * fb[(size_t)6 * HW]) + (T)m * depth_gt[idx];
}

}  // namespace warp

namespace scnet { extern long long g_conv_launches; }

extern "C" {

int rp_warp_workspace_bytes(int B, size_t* bytes) {
    if (B < 0 || !bytes) return RP_ERR_INVALID_ARG;
    *bytes = (size_t)B * warp::HW * sizeof(int);
    return RP_OK;
}

int rp_warp_views_ex(const float* view, long long view_img_stride, const int32_t* src_index, const double* R, int B, int dataset,
                     float* out, long long out_img_stride, void* workspace, size_t workspace_bytes, void* stream_) {
    if (B == 0) return RP_OK;
    if (!view || !R || !out || !workspace || B < 0 || dataset < 0 || dataset > 2) return RP_ERR_INVALID_ARG;
    if (view_img_stride < 8LL * warp::HW || out_img_stride < 8LL * warp::HW) return RP_ERR_INVALID_ARG;
    if (workspace_bytes < (size_t)B * warp::HW * sizeof(int)) return RP_ERR_WORKSPACE_TOO_SMALL;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    int* zbuf = static_cast<int*>(workspace);
    if (cudaMemsetAsync(zbuf, 0xFF, (size_t)B * warp::HW * sizeof(int), stream) != cudaSuccess) return RP_ERR_CUDA;
    const warp::Window wd = warp::window_of(dataset);
    const int ns = B * wd.w * wd.h, nt = B * warp::HW;
    warp::warp_scatter_kernel<<<(ns + 255) / 256, 256, 0, stream>>>(view, view_img_stride, src_index, R, B, dataset, zbuf);
    warp::warp_gather_kernel<<<(nt + 255) / 256, 256, 0, stream>>>(view, view_img_stride, src_index, R, B, dataset, zbuf, out, out_img_stride);
    scnet::g_conv_launches += 2;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_warp_views(const float* view, const double* R, int B, int dataset, float* out, void* workspace, size_t workspace_bytes,
                  void* stream_) {
    return rp_warp_views_ex(view, 8LL * warp::HW, nullptr, R, B, dataset, out, 8LL * warp::HW, workspace, workspace_bytes, stream_);
}

int rp_pano2pc(const float* depth, int B, int dataset, double* pc, unsigned char* valid, void* stream_) {
    if (B == 0) return RP_OK;
    if (!depth || !pc || B < 0 || dataset < 0 || dataset > 2) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const int n = B * warp::HW;
    warp::pano2pc_kernel<<<(n + 255) / 256, 256, 0, stream>>>(depth, B, dataset, pc, valid);
    ++scnet::g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

int rp_blend_completion(const float* f, int C, const float* mask, const void* norm_gt, const void* depth_gt, int is_f64, int B,
                        void* normal_out, void* depth_out, void* stream_) {
    if (B == 0) return RP_OK;
    if (!f || !mask || !norm_gt || !depth_gt || !normal_out || !depth_out || B < 0 || C < 7) return RP_ERR_INVALID_ARG;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const int n = B * warp::HW;
    if (is_f64)
        warp::blend_kernel<double><<<(n + 255) / 256, 256, 0, stream>>>(f, C, mask, static_cast<const double*>(norm_gt),
            static_cast<const double*>(depth_gt), B, static_cast<double*>(normal_out), static_cast<double*>(depth_out));
    else
        warp::blend_kernel<float><<<(n + 255) / 256, 256, 0, stream>>>(f, C, mask, static_cast<const float*>(norm_gt),
            static_cast<const float*>(depth_gt), B, static_cast<float*>(normal_out), static_cast<float*>(depth_out));
    ++scnet::g_conv_launches;
    return cudaGetLastError() == cudaSuccess ? RP_OK : RP_ERR_CUDA;
}

}  // extern "C"

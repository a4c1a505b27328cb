/* rp_b200.h -- C ABI of the B200-native relative-pose hot path.
 *
 * The reference (zhenpeiyang/RelativePose @ 2e9fdf5) has no FFI/plugin layer; its
 * boundary is a Python call surface.  This header is the
 * C-ABI a maintainer would bind underneath that surface (ctypes stub shown in
 * INTEGRATION.md).  Every entry point cites the reference function it replaces.
 *
 * Conventions: plain pointers and sizes only; all data pointers are DEVICE
 * pointers unless the name ends in _host; `stream` is a cudaStream_t passed as
 * void*; every function returns 0 on success or a negative RP_ERR_* code and
 * never throws; no allocation happens in the steady state (caller-owned
 * workspace, size from the *_workspace_bytes query); calls are asynchronous on
 * `stream` and re-entrant as long as concurrent calls use disjoint workspaces.
 *
 * Ragged batches: pair b owns source keypoints [off_s[b], off_s[b+1]) and target
 * keypoints [off_t[b], off_t[b+1]) of the concatenated keypoint arrays.
 */
#ifndef RP_B200_H
#define RP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RP_ABI_VERSION 1

/* error codes (function return values) */
#define RP_OK 0
#define RP_ERR_INVALID_ARG (-1)
#define RP_ERR_WORKSPACE_TOO_SMALL (-2)
#define RP_ERR_UNSUPPORTED (-3)   /* topK > RP_MAX_TOPK, feature dim > 128, n_t too large for shared memory */
#define RP_ERR_CUDA (-4)
#define RP_ERR_NO_DEVICE (-5)

/* per-pair status (status[b]); 1..5 are the reference's identity early exits */
#define RP_STATUS_OK 0
#define RP_STATUS_FEW_KEYPOINTS 1   /* rpmodule.py:346-348 */
#define RP_STATUS_FEW_CORRES 2      /* rpmodule.py:377-379 */
#define RP_STATUS_FEW_DIST 3        /* rpmodule.py:406-408 */
#define RP_STATUS_FEW_ANGLE 4       /* rpmodule.py:440-443 */
#define RP_STATUS_ZERO_WEIGHT 5     /* rpmodule.py:469-472 */
#define RP_STATUS_EDGE_OVERFLOW (-2) /* pair needs a larger edge capacity; re-run with a bigger workspace */
#define RP_STATUS_UNSUPPORTED (-3)

#define RP_MAX_TOPK 8
#define RP_MAX_FEAT_DIM 128

/* para.method (rputil.py:22, dispatch rpmodule.py:491-508) */
#define RP_METHOD_HORN87 0
#define RP_METHOD_SPECTRAL 1
#define RP_METHOD_IRLS 2
#define RP_METHOD_IRLS_SM 3

/* stages for rp_solve_batch_ex(stop_after) */
#define RP_STAGE_TOPK 1      /* rpmodule.py:342-375 */
#define RP_STAGE_AFFINITY 2  /* rpmodule.py:382-472 */
#define RP_STAGE_SOLVE 3     /* rpmodule.py:484-508 */

/* Solver parameters: `rputil.opts` (RPModule/rputil.py:11-22) reduced on the HOST, with
 * numpy arithmetic, to the scalars the reference actually uses, so thresholds are the
 * reference's bit patterns:
 *   feat_den      = 2*np.power(sigmaFeat/5, 2)              rpmodule.py:356,358
 *   feat_den_obs  = 2*np.power((sigmaFeat/1.2)/5, 2)        rpmodule.py:357,358
 *   dist_thre_sq  = np.power(distThre, 2)                   rpmodule.py:404
 *   sep_thre      = 1.5*np.power(distSepThre, 2)            rpmodule.py:404
 *   angle_thre_sq = np.power(angleThre, 2)                  rpmodule.py:434-436
 *   den_dist      = 2*sigmaDist**2, den_a1 = 2*sigmaAngle1**2, den_a2 = 2*sigmaAngle2**2   rpmodule.py:457-460
 */
typedef struct rp_params {
    double feat_den;
    double feat_den_obs;
    double dist_thre_sq;
    double sep_thre;
    double angle_thre_sq;
    double den_dist;
    double den_a1;
    double den_a2;
    double mu;          /* rputil.py:20 */
    double power_tol;   /* stop power iteration when ||u_k - u_{k-1}||_2 <= power_tol (reference: ARPACK tol=0) */
    int32_t topk;       /* rputil.py:21; effective K = min(topk, n_t-1), rpmodule.py:368 */
    int32_t method;     /* RP_METHOD_* */
    int32_t max_power_iters;
    int32_t reserved;
} rp_params;

#define RP_STATS_STRIDE 8
/* stats[b*8 + k]: 0 N (=n_s*K), 1 pairs after distance test, 2 pairs after angle test,
 * 3 pairs with non-zero weight, 4 power iterations (sum over alternations),
 * 5 max power iterations in one alternation, 6 bit 0: some alternation hit max_power_iters; bits 8..: number of source
 * rows whose top-k index SET the keys alone do not determine (the K-th and (K+1)-th candidate tie exactly, or selected
 * entries underflowed to weight 0 in a row that is not all zero): there the reference's numpy.argpartition order decides and
 * the sets may differ from it (all-zero rows follow numpy through zero_row_topk; counted on the 32-channel path),
 * 7 bits 0..7: K; bit 8: the pair was solved by the accelerated (locally optimal CG) eigen iteration */

/* Optional stage-boundary outputs (device pointers; any may be NULL).  Used by the parity tests. */
typedef struct rp_debug {
    int32_t* topk_idx;   /* [sum n_s, max_topk] target index of each candidate, -1 padded; order = descending weight */
    double* topk_f;      /* [sum n_s, max_topk] normalised weight wij of each candidate (rpmodule.py:362,453-454) */
    float* dij;          /* per pair n_s*n_t float32 descriptor distances at dij_off[b] (rpmodule.py:355) */
    const int64_t* dij_off; /* [B] element offsets into dij */
    int32_t* edge_rc;    /* [B, edge_cap, 2] surviving (first,second) correspondence ids, row-major order NOT guaranteed */
    double* edge_w;      /* [B, edge_cap] pair weight w (rpmodule.py:457-467) */
    int64_t edge_cap;
    double* u;           /* [B, 5, u_stride] leading eigenvector per alternation (irls+sm / spectral) */
    int64_t u_stride;
    int64_t* phase_clk;  /* [B, 8] SM clock (clock64) of the pair's CTA at: 0 start, 1 after the descriptor front end, 2 after the
                          * pre-test, 3 after the exact tests, 4 after the CSR build, 5 end; 6 cycles inside the eigen iterations,
                          * 7 cycles inside the Horn fits + residual passes (profiling aid) */
} rp_debug;

int rp_abi_version(void);

/* 16-bit operand / activation-storage format of the tensor-core convolution kernels: 1 = IEEE half (default build),
 * 0 = bfloat16.  Buffers with rp_conv_src.dtype == 1 / out_dtype == 1 hold this format. */
int rp_h16_format(void);

/* Number of SMs / device name of the current CUDA device (host utility). */
int rp_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* Workspace for rp_solve_batch*: `n_slots` resident CTAs (0 = default: one or two per SM),
 * each able to hold a pair with n_s <= max_ns, n_t <= max_nt, K <= max_topk and up to
 * `edge_cap` pairs passing the distance test (0 = worst case N(N-1)/2). */
int rp_solve_workspace_bytes(int n_slots, int max_ns, int max_nt, int max_topk, int feat_dim,
                             int64_t edge_cap, size_t* bytes);

/* The default slot count (resident CTAs on the current device) for pairs of this size: callers with fewer than that
 * many pairs size the workspace for min(B, default) slots.  Pairs with n_s*topK <= 16383 are supported; above ~2800
 * correspondences (or n_t > ~1400) the per-pair vectors move from shared memory into the slot's workspace. */
int rp_solve_default_slots(int max_ns, int max_nt, int max_topk, int feat_dim, int* n_slots);

/* One scan pair from HOST arrays (the reference's call pattern: evaluation.py:278-284 calls RelativePoseEstimation_helper,
 * RPModule/rpmodule.py:317-508, once per pair).  All pointers are host pointers: pc / nrm [n,3] float64, feat [n,feat_dim]
 * float32 (C order), w [n] float64; params: one rp_params; zero_row_topk: max_topk ints (numpy's tie order for all-zero rows)
 * or NULL; feat_sum_order as in rp_solve_batch; edge_cap: candidate capacity (0 = worst case).  Packs the inputs into one
 * page-locked block, ONE host-to-device copy, the fused launch, ONE device-to-host copy of T_out[16], status[1], stats[8]
 * (stats may be NULL), and synchronises `stream`.  Device staging and workspace are cached per device inside the library.
 * RP_STATUS_EDGE_OVERFLOW in *status: call again with edge_cap = 0. */
int rp_solve_pair_host(int ns, int nt,
                       const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                       const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                       int feat_dim, const rp_params* params, const int32_t* zero_row_topk, int max_topk,
                       int feat_sum_order, int64_t edge_cap,
                       double* T_out, int32_t* status, int32_t* stats, void* stream);

/* Small-batch path.  The solver kernel is built twice: 128-thread CTAs, four per SM (throughput for batches that fill the
 * GPU) and 512-thread CTAs, one per SM, which give a scan pair a whole SM.  The reference calls its solver one pair at a time
 * (evaluation.py:278-284) and the completion alternation solves a few dozen pairs per step (rpmodule.py:569-662); batches of
 * at most `wide_max` pairs take the 512-thread build (default: the SM count; environment RP_SOLVER_WIDE_MAX; 0 = never).
 * rp_solver_wide_max(new_max) sets the limit when new_max >= 0 and returns the previous one.  The two builds agree to
 * rounding (reduction trees differ with the CTA width), not bitwise; which build a batch takes depends on its size alone. */
int rp_solver_wide_max(int new_max);

/* Replaces RelativePoseEstimation_helper (RPModule/rpmodule.py:317-508) for a ragged batch of B
 * scan pairs -- descriptor distance + soft match + top-k, pairwise consistency affinity,
 * spectral/IRLS solve -- one fused launch.
 *   pc_*   [sum n, 3] float64   'pc'      nrm_* [sum n, 3] float64 'normal'
 *   feat_* [sum n, feat_dim] float32 'feat' (before /100)   w_* [sum n] float64 'weight'
 *   params [n_params] (device), param_idx [B] (device) or NULL (all pairs use params[0])
 *   zero_row_topk [B, max_topk] (device) or NULL: the candidate set to use for a source keypoint whose
 *     soft-match row underflows to all zeros (rpmodule.py:359-363 zeroes the row; np.argpartition then
 *     returns an order-of-introselect set that depends only on (n_t, K)).  The host fills it with
 *     np.argpartition(-np.zeros(n_t), K)[:K] so ties resolve exactly like the reference; NULL = by distance.
 *   feat_sum_order [B] (device) or NULL (= all 0): float32 summation order of the descriptor distance, which in
 *     NumPy depends on the memory layout of the 'feat' arrays (rpmodule.py:355): 0 = both C-contiguous (8-lane
 *     pairwise sum), 1 = either one a transposed view, as the reference's own pipeline passes them
 *     (rpmodule.py:531-532) -> sequential sum over the channels.  The Python layer derives it from the array flags.
 *   T_out [B,16] row-major 4x4 float64; status [B]; stats [B,8] (may be NULL)
 */
int rp_solve_batch(int B, const int32_t* off_s, const int32_t* off_t,
                   const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                   const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                   int feat_dim, const rp_params* params, const int32_t* param_idx,
                   const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                   int max_ns, int max_nt, int max_topk,
                   int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                   double* T_out, int32_t* status, int32_t* stats, void* stream);

/* Same, stopping after a stage (RP_STAGE_*) and filling the stage-boundary outputs in `dbg`
 * (host struct holding device pointers).  rp_match_topk / rp_affinity_build are this call with
 * stop_after = RP_STAGE_TOPK / RP_STAGE_AFFINITY. */
int rp_solve_batch_ex(int B, const int32_t* off_s, const int32_t* off_t,
                      const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                      const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                      int feat_dim, const rp_params* params, const int32_t* param_idx,
                      const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                      int max_ns, int max_nt, int max_topk,
                      int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                      double* T_out, int32_t* status, int32_t* stats,
                      int stop_after, const rp_debug* dbg, void* stream);

/* Stage entry: rpmodule.py:342-472 (descriptor weights, top-k, pair enumeration, distance / angle filters, pair weight):
 * the geometrically consistent pairs of every scan pair.  topk_idx [sum n_s, max_topk] (-1 padded), edge_rc
 * [B, edge_cap, 2] = the two correspondence indices (source index * K + rank) of each surviving pair, edge_w [B, edge_cap];
 * the number of pairs is stats[b*8 + 2].  `edge_cap` is the capacity of the two output arrays; the workspace is the one of
 * rp_solve_batch with its edge_cap = 0 (worst case). */
int rp_affinity_build(int B, const int32_t* off_s, const int32_t* off_t,
                      const double* pc_s, const double* nrm_s, const float* feat_s, const double* w_s,
                      const double* pc_t, const double* nrm_t, const float* feat_t, const double* w_t,
                      int feat_dim, const rp_params* params, const int32_t* param_idx,
                      const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                      int max_ns, int max_nt, int max_topk,
                      int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                      int32_t* topk_idx, int32_t* edge_rc, double* edge_w, int32_t* status, int32_t* stats, void* stream);

/* Stage entry: rpmodule.py:342-375 only (dij -> wij -> row-normalise -> top-k). */
int rp_match_topk(int B, const int32_t* off_s, const int32_t* off_t,
                  const float* feat_s, const double* w_s, const float* feat_t, const double* w_t,
                  int feat_dim, const rp_params* params, const int32_t* param_idx,
                  const int32_t* zero_row_topk, const int32_t* feat_sum_order,
                  int max_ns, int max_nt, int max_topk,
                  int n_slots, void* workspace, size_t workspace_bytes,
                  int32_t* topk_idx, double* topk_f, int32_t* status, void* stream);


/* ------------------------------------------------------------------------------------------------------------
 * SCNet / Resnet18_8s layer ABI (model/mymodel.py).  The reference runs these through torch.nn / cuDNN
 * (nn.Conv2d, nn.ConvTranspose2d, BatchNorm2d(track_running_stats=False), LeakyReLU(0.1), F.upsample); here one
 * kernel family does "conv block" = [normalise + LeakyReLU of the producer's raw output while loading] ->
 * conv / transposed conv -> raw output + per-(pair, channel) partial batch statistics.  Activations are float32
 * NHWC with an explicit channel pitch/offset, so torch.cat (mymodel.py:266-325) never materialises.
 * A "group" is one scan pair = 2 consecutive images = one forward call of the reference = one BN batch.
 */
typedef struct rp_conv_src {
    const float* ptr;     /* [2G, Hin, Win, pitch] raw (pre-BN) activations */
    int32_t pitch;        /* channels per pixel in memory */
    int32_t ch_off;       /* first channel of this view */
    int32_t C;            /* channels used */
    int32_t act;          /* 0: use as is; 1: x*scale+shift then LeakyReLU(0.1) (mymodel.py:19-20,32-33) */
    const float* scale;   /* [G, sstride] per-(group, channel) BN scale  gamma/sqrt(var+eps)   (act=1) */
    const float* shift;   /* [G, sstride] per-(group, channel) BN shift  beta - mean*scale */
    int32_t sstride;
    int32_t s_off;
    float slope;          /* negative slope of the activation applied with act=1: 0.1 = LeakyReLU (SCNet), 0 = ReLU (ResNet) */
    int32_t dtype;        /* storage type of ptr: 0 = float32, 1 = bfloat16 (tensor-core kernels only) */
} rp_conv_src;

typedef struct rp_conv_desc {
    rp_conv_src src[2];   /* channel-concatenated inputs (torch.cat((a,b),1)) */
    int32_t nsrc;
    int32_t transposed;   /* 0 nn.Conv2d, 1 nn.ConvTranspose2d */
    int32_t k, s, p;      /* kernel, stride, padding (square) */
    int32_t G;            /* scan pairs */
    int32_t Hin, Win, Hout, Wout;
    int32_t Cout;
    const float* W;       /* packed [k*k][Cin_total][Cout]: conv w[co,ci,ky,kx], transposed w[ci,co,ky,kx] */
    float* out;           /* [2G, Hout, Wout, out_pitch] raw output */
    int32_t out_pitch, out_ch_off;
    float* psum;          /* [G, nparts, Cout] partial sums for the batch statistics, or NULL */
    float* psq;
    const float* bias;    /* [Cout] or NULL (the 1x1 heads, mymodel.py:188,196,204,220,228) */
    int32_t tanh_out;     /* mymodel.py:374-375 */
    int32_t imgs_per_group; /* images per BN batch: 0 or 2 = one scan pair (SCNet); Resnet18_8s uses the whole call */
    int32_t out_dtype;    /* storage type of out: 0 = float32, 1 = bfloat16 (tensor-core kernels and the small-Cin stem) */
} rp_conv_desc;

/* number of partial-statistics rows per group the layer writes (psum/psq are [G, nparts, Cout]) */
int rp_conv_nparts(const rp_conv_desc* d, int* nparts);
int rp_conv_layer(const rp_conv_desc* d, void* stream);
/* scale/shift [G, sstride] (+s_off) from the partials: batch mean / biased variance over the 2*Hout*Wout
 * pixels of a group (`count`), eps 1e-5 (nn.BatchNorm2d, mymodel.py:19,32) */
int rp_bn_finalize(const float* psum, const float* psq, int G, int nparts, int Cout, int count,
                   const float* gamma, const float* beta, float* scale, float* shift, int sstride, int s_off,
                   void* stream);
/* same for very long partial lists (Resnet18_8s: the BatchNorm batch is the whole call): nsplit row slices summed in
 * float64 into scratch [G, nsplit, Cout, 2] doubles, then combined in order */
int rp_bn_finalize_split(const float* psum, const float* psq, int G, int nparts, int Cout, int count,
                         const float* gamma, const float* beta, float* scale, float* shift, int sstride, int s_off,
                         int nsplit, double* scratch, void* stream);
/* im2col of NHWC float32 [n,H,W,C] into bfloat16 rows [n,Hout,Wout,Kpad] (K = (ky*k+kx)*C + c, zero padded): the 7x7/s2
 * stem of Resnet18_8s (mymodel.py:51-54,85) becomes a 1x1 convolution with K = 352 on the tensor cores */
int rp_im2col_bf16(const float* x, int n, int H, int W, int C, int k, int s, int p, int Hout, int Wout, int Kpad, void* out, void* stream);
/* Space-to-depth of an NCHW float32 input [n,C,H,W] (H, W even, 4 C <= Cpad, Cpad % 8 == 0) into 16-bit NHWC [n,H/2,W/2,Cpad]:
 * channel (dy*2+dx)*C + c of pixel (sy,sx) = x[c, 2 sy + dy, 2 sx + dx], the rest zero.  A k = 7, stride 2, padding 3 convolution
 * (the Resnet18_8s stem, mymodel.py:51-54,85) is then a 4x4 stride-1 convolution with padding 2 over this tensor (weights
 * W4[by,bx,(dy*2+dx)*C+c] = W7[2 by + dy - 1, 2 bx + dx - 1, c], zero where the index leaves the 7x7 kernel) -- 16 taps over 32
 * channels on the halo kernel instead of an im2col matrix 10x the size of the input. */
int rp_space_to_depth_h16(const float* x, int n, int C, int H, int W, int Cpad, void* out, void* stream);
/* F.upsample(x,[224,224],'bilinear',align_corners=False) (mymodel.py:261) fused with the channel regrouping of
 * mymodel.py:264-286: in [n,16,H,W] NCHW -> out [n,224,224,20] NHWC = (rgb,mask | normal,mask | depth,mask) x (own, warped) */
int rp_scnet_resize_in(const float* x, int n, int H, int W, float* out, void* stream);
/* Same for the tensor-core stem: out [n,224,224,96] bfloat16 = 6 groups x [hi(4) | lo(4) | hi(4) | 0(4)], x = hi + lo
 * (bf16 split), so that conv1* runs on tcgen05 with float32-class accuracy (csrc/scnet.cu). */
int rp_scnet_resize_in_split(const float* x, int n, int H, int W, void* out, void* stream);
/* F.upsample(xout,inShape,'bilinear',align_corners=False) (mymodel.py:379): in [n,224,224,C] NHWC -> out [n,C,H,W] NCHW */
int rp_scnet_resize_out(const float* in, int n, int C, int H, int W, float* out, void* stream);
/* same with `pitch` floats per pixel in memory and a device channel map (output channel c reads cmap[c]): the engine keeps
 * every head at a 16-byte aligned channel offset of the 224x224 tensor */
int rp_scnet_resize_out_map(const float* in, int n, int pitch, const int* cmap, int C, int H, int W, float* out, void* stream);
/* the same for a subset of the channels of an output tensor [n, out_channels, H, W]: item i reads source channel cmap[i] and writes
 * output channel omap[i] (device arrays of C ints); the other output channels are not touched.  The batched alternation
 * (relativepose_b200/pipeline.py) only reads the normal / depth / descriptor heads of the completion network, so the other heads
 * are neither computed nor resized there */
int rp_scnet_resize_out_sub(const float* in, int n, int pitch, const int* cmap, const int* omap, int C, int H, int W, float* out,
                            int out_channels, void* stream);

/* Stage entry: the fitters only (rpmodule.py:484-508; fit_horn87 :60, fit_spectral :86, fit_irls :169, fit_irls_sm :212,
 * horn87_np :17).  Problem b owns nodes [node_off[b], node_off[b+1]) = candidate correspondences with source/target
 * position (sp, tp) and normal (sn, tn), and optionally edges [edge_off[b], edge_off[b+1]) = geometrically consistent
 * pairs: two local node indices and the pair weight w (rpmodule.py:457-467).  With edges the base weights are the row
 * degrees of W (exactly the reference's stacked rows allWP = allWN = [w, w]); with edge_off == NULL the explicit
 * per-node weights node_wp/node_wn are used (allWP/allWN of fit_horn87 / fit_irls; horn87_np = normals only). */
int rp_spectral_irls_workspace_bytes(int n_slots, int max_nodes, int64_t edge_cap, size_t* bytes);
int rp_spectral_irls_solve(int B, const int32_t* node_off,
                           const double* sp, const double* sn, const double* tp, const double* tn,
                           const double* node_wp, const double* node_wn,
                           const int32_t* edge_off, const int32_t* edge_rc, const double* edge_w,
                           const rp_params* params, const int32_t* param_idx, int max_nodes,
                           int n_slots, int64_t edge_cap, void* workspace, size_t workspace_bytes,
                           double* T_out, int32_t* status, int32_t* stats, void* stream);

/* Kernel launch counter (number of kernels this library launched since load); bench.py reports it. */
/* ---- Resnet18_8s extras (mymodel.py:82-122; stock ResNet-18 trunk) -------------------------------------------
 * y = act(x*scale+shift) with per-(group,channel) BN scale/shift (NULL = identity), all tensors float32 NHWC. */
/* relu(bn1(conv1)) followed by MaxPool2d(3, stride 2, padding 1): in [n,H,W,C] raw -> out [n,Ho,Wo,C] activated */
int rp_bn_relu_maxpool(const float* in, int n, int H, int W, int C, int imgs_per_group,
                       const float* scale, const float* shift, float* out, int Ho, int Wo, void* stream);
/* BasicBlock tail: out = relu(a*sa+ha + (b*sb+hb)); sb/hb NULL = identity shortcut (b already activated) */
int rp_bn_add_relu(const float* a, const float* sa, const float* ha, const float* b, const float* sb, const float* hb,
                   float* out, int n, int HW, int C, int imgs_per_group, void* stream);
/* F.upsample(src, size, 'bilinear', align_corners=False) on NHWC: dst = (accumulate ? dst : 0) + up(src) */
int rp_resize_nhwc(const float* src, int n, int Hs, int Ws, int C, float* dst, int Hd, int Wd, int accumulate, void* stream);
/* final F.upsample to the input size + optional tanh (tanh_out 1 = tanhf, 2 = tanh.approx.f32 for the bf16 tensor-core mode),
 * NHWC [n,Hs,Ws,C] -> NCHW [n,C,H,W] (mymodel.py:111,120-121) */
int rp_resize_to_nchw(const float* src, int n, int Hs, int Ws, int C, float* out, int H, int W, int tanh_out, void* stream);

/* rputil.interpolate (RPModule/rputil.py:43-58): feat [C,H,W], pt [K,2] normalised (x,y) -> out [C,K]; device float32 */
int rp_interpolate(const float* feat, int C, int H, int W, const float* pt, int K, float* out, void* stream);

/* ---- View warping between the two scans of a pair --------------------------------------
 * util.warping (util.py:94-172) over a batch: view [B,8,160,640] float32 NCHW (rgb, normal, depth, valid), R [B,16]
 * row-major 4x4 float64, dataset 0 = suncg, 1 = matterport, 2 = scannet -> out [B,8,160,640] float32 (the reference's
 * float64 result cast the way its caller does, torch_op.v).  The observed window of the view (skybox face 1 / the
 * 66x88 Kinect window) is lifted to 3-D (util.depth2pc / Pano2PointCloud), moved by R and splatted on the 4 faces
 * (util.reproj_helper: numpy "last write wins" = the source pixel with the largest raster index wins).  An identity R
 * yields zeros (util.py:95-96).  workspace: rp_warp_workspace_bytes(B) bytes (the int32 winner map). */
int rp_warp_workspace_bytes(int B, size_t* bytes);
int rp_warp_views(const float* view, const double* R, int B, int dataset, float* out, void* workspace, size_t workspace_bytes,
                  void* stream);
/* same with explicit image strides (floats) and an optional source index: output b = view[src_index[b]] warped by R[b], written
 * at out + b*out_img_stride -- the batched alternation warps every scan's partner straight into channels 8..15 of the network
 * input (rpmodule.py:616-621) without materialising views[swap] or the torch.cat */
int rp_warp_views_ex(const float* view, long long view_img_stride, const int32_t* src_index, const double* R, int B, int dataset,
                     float* out, long long out_img_stride, void* workspace, size_t workspace_bytes, void* stream);
/* util.Pano2PointCloud (util.py:751-811), dense: depth [B,160,640] float32 -> pc [B,3,102400] float64 in the reference's
 * point order (face, row, column); valid [B,102400] (may be NULL) = 1 where the reference keeps the point (scannet drops
 * depth == 0, the caller compacts). */
int rp_pano2pc(const float* depth, int B, int dataset, double* pc, unsigned char* valid, void* stream);
/* The blend of RelativePoseEstimationViaCompletion (rpmodule.py:628-634): f [B,C,160,640] float32 network output (3:6
 * normal, 6 depth), mask [B,160,640] float32, norm_gt [B,160,640,3] / depth_gt [B,160,640] float64 (is_f64) or float32
 * -> normal_out / depth_out of the same type: observed region from the scan, the rest from the network, normals
 * re-normalised (EPS 1e-12). */
int rp_blend_completion(const float* f, int C, const float* mask, const void* norm_gt, const void* depth_gt, int is_f64, int B,
                        void* normal_out, void* depth_out, void* stream);

/* ---- Keypoint augmentation (RPModule/rputil.py:179-214 + Sampling :355-371) ----------------
 * For every query descriptor q[:, i] (q [C, nq] float32, the layout rputil.interpolate returns) the squared distance to
 * every pixel of feat [C,H,W] is formed and K rounds of { argmax of exp(-d/2); suppress [y-window, min(H-1,y+window)) x
 * [x-window, min(W-1,x+window)) with the map's minimum } pick pts [nq, K, 2] float64 (x, y).  rp_heat_sample is `Sampling`
 * alone on a caller-provided distance map dist [n,H,W].  Device pointers. */
int rp_match_sample_workspace_bytes(int nq, size_t* bytes);
int rp_match_sample(const float* q, int C, int nq, const float* feat, int H, int W, int K, int window, double* pts,
                    void* workspace, size_t workspace_bytes, void* stream);
int rp_heat_sample(const float* dist, int n, int H, int W, int K, int window, double* pts,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Net -> solver hand-off on the device (rpmodule.getMatchingPrimitive, rpmodule.py:511-538, fixed keypoint counts): for
 * image i and keypoint k at pixel pts[i,k] = (x,y) (float64): rputil.getPixel (rputil.py:61-119) on depth [n_img,160,640] /
 * normal [n_img,160,640,3] (float64) -> pc_out / nn_out [n_img,K,3]; rputil.interpolate (rputil.py:43-58) of the C-channel
 * float32 map feat + i*feat_img_stride ([C,160,640]) -> desc_out [n_img,K,C] (the row layout rp_solve_batch reads). */
int rp_gather_primitives(const float* feat, int C, long long feat_img_stride, const double* depth, const double* normal,
                         const double* pts, int n_img, int K, int dataset, double* pc_out, double* nn_out, float* desc_out,
                         void* stream);

/* ---- A network forward as one native call (scnet_forward / resnet18_8s_forward) -----------------
 * The host freezes the layer calls of one forward into an op list once per (input shape, weights): op i = one of the layer
 * entry points above (`kind`) with its rp_conv_desc and/or up to 16 scalar arguments in call order (pointers and integers as
 * 64-bit words).  rp_scnet_forward (model/mymodel.py:259-380) / rp_resnet18_8s_forward (:82-122) issue every launch on
 * `stream`; the buffers the ops point to must stay allocated (relativepose_b200/scnet_engine.py keeps them per shape). */
enum { RP_OP_CONV = 1, RP_OP_CONV_TC_REMOVED /* 2: per-tap tcgen05 kernel, removed */, RP_OP_CONV_HALO, RP_OP_BN_FINALIZE, RP_OP_BN_FINALIZE_SPLIT, RP_OP_RESIZE_IN,
       RP_OP_RESIZE_IN_SPLIT, RP_OP_RESIZE_OUT_MAP, RP_OP_IM2COL, RP_OP_BN_RELU_MAXPOOL, RP_OP_BN_ADD_RELU, RP_OP_RESIZE_NHWC,
       RP_OP_RESIZE_TO_NCHW, RP_OP_SPACE_TO_DEPTH, RP_OP_RESIZE_OUT_SUB };
typedef struct rp_net_op {
    int32_t kind;         /* RP_OP_* */
    int32_t reserved;
    rp_conv_desc conv;    /* RP_OP_CONV / _TC / _HALO */
    uint64_t arg[16];     /* the remaining arguments of the call, in order (without the trailing stream) */
} rp_net_op;
int rp_scnet_forward(const rp_net_op* ops, int n_ops, void* stream);
int rp_resnet18_8s_forward(const rp_net_op* ops, int n_ops, void* stream);

int64_t rp_launch_count(void);
int64_t rp_conv_launch_count(void);

/* Halo-tile variant (csrc/scnet_halo.cu): a CTA stages the input halo of a 16x8 block of output positions once per
 * K chunk and every (sub-pixel class, tap) MMA reads it through a shifted shared-memory descriptor, so a 4x4 kernel no
 * longer re-gathers its input 16 times.  bn in {32,64,128} with Cout % bn == 0, tk in {32,64} with every source
 * C % tk == 0, at most 16 (class, tap) blocks, no bias/tanh.  flags bit 0: pad the halo row pitch to 16 pixels.
 * rp_conv_halo_plan returns the partial-statistics rows per group, the number of (class, tap) weight blocks per
 * K chunk and their order (tap_widx[i] = ky*k+kx, room for 16); w_packed is bf16
 * [n-tile][K chunk][block][tk/8][bn/8][8 co][8 ci] (relativepose_b200/scnet_engine.py:pack_halo). */
int rp_conv_halo_plan(const rp_conv_desc* d, int bn, int tk, int flags, int* nparts, int* ntap, int* tap_widx);
/* Launches of the halo kernel so far whose input halo was fetched by tiled TMA (16-bit sources; float32 sources use cp.async /
 * thread gathers). */
long long rp_conv_halo_tma_count(void);

/* Profiling aid: per-role cycle counters of the halo kernel (enabled by flags bit 5 of rp_conv_layer_halo), read and cleared. */
int rp_conv_halo_prof(unsigned long long* out16);

/* RP_OK when the tile plan of this layer fits one CTA's shared memory under `flags` (the fused split-precision launch, flags bit
 * 10, keeps a hi and a lo halo per buffer), RP_ERR_UNSUPPORTED otherwise.  Launches nothing. */
int rp_conv_halo_fits(const rp_conv_desc* d, int bn, int tk, int flags);

/* flags of rp_conv_layer_halo: bit 0 16-pixel halo pitch; bit 1 packed-half BatchNorm of 16-bit sources; bits 2-5 profiling /
 * debugging; bit 6 / 7 never / always tiled TMA for 16-bit sources.  Split-precision launches (float32 sources and output; the
 * network mode RP_SCNET_MODE=tc3) compute x w = half(x') hi(w') + lo(x') hi(w') + half(x') lo(w') with x' = 2^4 x, w' = 2^8 w,
 * lo(v) = v - half(v), the epilogue scaling by 2^-12:  bit 10 all three terms in ONE launch -- w_packed holds (hi(w'), lo(w'))
 * block pairs, the loader fills a hi and a lo halo, three tcgen05.mma per tap and K step accumulate into one TMEM accumulator
 * (needs room for the doubled halo: rp_conv_halo_fits);  otherwise THREE launches, each term in an accumulator of its own: the
 * plain launch half(x) hi(w) stores, then bits 8 + 9 (the loader emits lo(x'), w_packed = the hi(w') blocks) and bit 9 (w_packed =
 * the lo(w') blocks) add 2^-12 x their accumulators onto the stored float32 output; bias / tanh / statistics go with the last.
 * (Bit 11: block pairs over a single halo, half(x') [hi(w') + lo(w')] in one launch -- measurably less accurate, unused.) */
int rp_conv_layer_halo(const rp_conv_desc* d, const void* w_packed, int bn, int tk, int flags, void* stream);
/* test hook: the tile plan as 96 integers (layout in csrc/scnet_halo.cu) for the CPU emulation in tests/test_halo_plan.py */
int rp_conv_halo_debug(const rp_conv_desc* d, int bn, int tk, int flags, int* out96);

/* Unit-test hook for the tcgen05/TMEM building blocks: C[M,N] = A[M,K] * B[N,K]^T (device pointers, float32 in/out,
 * bf16-rounded operands, fp32 accumulation in TMEM).  M % 128 == 0, K % 64 == 0, bn in {64,128}, N % bn == 0. */
int rp_tc_gemm_test(const float* A, const float* B, float* C, int M, int N, int K, int bn, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RP_B200_H */

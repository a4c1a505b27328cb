// rp_plan.cu -- a network forward as ONE native call (`scnet_forward`, `resnet18_8s_forward`).
//
// The host builds the list of layer operations once per (input shape, weights) -- every entry is one of the layer calls
// of include/rp_b200.h with its arguments frozen: activation buffers, packed weights and BatchNorm parameters are
// persistent device allocations -- and hands it to rp_scnet_forward / rp_resnet18_8s_forward, which issue all launches
// on the stream.  No Python between the ~90 launches of SCNet.forward; the call is also what gets captured into a CUDA
// graph.  The reference's counterpart is nn.Module.__call__ walking its children (model/mymodel.py:259-380, :82-122).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rp_b200.h"

namespace {

template <typename T> inline T as_ptr(uint64_t v) { return reinterpret_cast<T>(static_cast<uintptr_t>(v)); }
inline int as_int(uint64_t v) { return static_cast<int>(static_cast<int64_t>(v)); }

int run_plan(const rp_net_op* ops, int n_ops, void* stream) {
    if (n_ops < 0 || (n_ops > 0 && !ops)) return RP_ERR_INVALID_ARG;
    for (int i = 0; i < n_ops; ++i) {
        const rp_net_op& o = ops[i];
        const uint64_t* a = o.arg;
        int rc = RP_ERR_INVALID_ARG;
        switch (o.kind) {
        case RP_OP_CONV: rc = rp_conv_layer(&o.conv, stream); break;
        case RP_OP_CONV_HALO: rc = rp_conv_layer_halo(&o.conv, as_ptr<const void*>(a[0]), as_int(a[1]), as_int(a[2]), as_int(a[3]), stream); break;
        case RP_OP_BN_FINALIZE:
            rc = rp_bn_finalize(as_ptr<const float*>(a[0]), as_ptr<const float*>(a[1]), as_int(a[2]), as_int(a[3]), as_int(a[4]), as_int(a[5]),
                                as_ptr<const float*>(a[6]), as_ptr<const float*>(a[7]), as_ptr<float*>(a[8]), as_ptr<float*>(a[9]),
                                as_int(a[10]), as_int(a[11]), stream);
            break;
        case RP_OP_BN_FINALIZE_SPLIT:
            rc = rp_bn_finalize_split(as_ptr<const float*>(a[0]), as_ptr<const float*>(a[1]), as_int(a[2]), as_int(a[3]), as_int(a[4]), as_int(a[5]),
                                      as_ptr<const float*>(a[6]), as_ptr<const float*>(a[7]), as_ptr<float*>(a[8]), as_ptr<float*>(a[9]),
                                      as_int(a[10]), as_int(a[11]), as_int(a[12]), as_ptr<double*>(a[13]), stream);
            break;
        case RP_OP_RESIZE_IN: rc = rp_scnet_resize_in(as_ptr<const float*>(a[0]), as_int(a[1]), as_int(a[2]), as_int(a[3]), as_ptr<float*>(a[4]), stream); break;
        case RP_OP_RESIZE_IN_SPLIT: rc = rp_scnet_resize_in_split(as_ptr<const float*>(a[0]), as_int(a[1]), as_int(a[2]), as_int(a[3]), as_ptr<void*>(a[4]), stream); break;
        case RP_OP_RESIZE_OUT_MAP:
            rc = rp_scnet_resize_out_map(as_ptr<const float*>(a[0]), as_int(a[1]), as_int(a[2]), as_ptr<const int*>(a[3]), as_int(a[4]), as_int(a[5]),
                                         as_int(a[6]), as_ptr<float*>(a[7]), stream);
            break;
        case RP_OP_IM2COL:
            rc = rp_im2col_bf16(as_ptr<const float*>(a[0]), as_int(a[1]), as_int(a[2]), as_int(a[3]), as_int(a[4]), as_int(a[5]), as_int(a[6]),
                                as_int(a[7]), as_int(a[8]), as_int(a[9]), as_int(a[10]), as_ptr<void*>(a[11]), stream);
            break;
        case RP_OP_BN_RELU_MAXPOOL:
            rc = rp_bn_relu_maxpool(as_ptr<const float*>(a[0]), as_int(a[1]), as_int(a[2]), as_int(a[3]), as_int(a[4]), as_int(a[5]),
                                    as_ptr<const float*>(a[6]), as_ptr<const float*>(a[7]), as_ptr<float*>(a[8]), as_int(a[9]), as_int(a[10]), stream);
            break;
        case RP_OP_BN_ADD_RELU:
            rc = rp_bn_add_relu(as_ptr<const float*>(a[0]), as_ptr<const float*>(a[1]), as_ptr<const float*>(a[2]), as_ptr<const float*>(a[3]),
                                as_ptr<const float*>(a[4]), as_ptr<const float*>(a[5]), as_ptr<float*>(a[6]), as_int(a[7]), as_int(a[8]),
                                as_int(a[9]), as_int(a[10]), stream);
            break;
        case RP_OP_RESIZE_NHWC:
            rc = rp_resize_nhwc(as_ptr<const float*>(a[0]), as_int(a[1]), as_int(a[2]), as_int(a[3]), as_int(a[4]), as_ptr<float*>(a[5]),
                                as_int(a[6]), as_int(a[7]), as_int(a[8]), stream);
            break;
        case RP_OP_RESIZE_TO_NCHW:
            rc = rp_resize_to_nchw(as_ptr<const float*>(a[0]), as_int(a[1]), as_int(a[2]), as_int(a[3]), as_int(a[4]), as_ptr<float*>(a[5]),
                                   as_int(a[6]), as_int(a[7]), as_int(a[8]), stream);
            break;
        case RP_OP_SPACE_TO_DEPTH:
            rc = rp_space_to_depth_h16(as_ptr<const float*>(a[0]), as_int(a[1]), as_int(a[2]), as_int(a[3]), as_int(a[4]), as_int(a[5]),
                                       as_ptr<void*>(a[6]), stream);
            break;
        case RP_OP_RESIZE_OUT_SUB:
            rc = rp_scnet_resize_out_sub(as_ptr<const float*>(a[0]), as_int(a[1]), as_int(a[2]), as_ptr<const int*>(a[3]), as_ptr<const int*>(a[4]),
                                         as_int(a[5]), as_int(a[6]), as_int(a[7]), as_ptr<float*>(a[8]), as_int(a[9]), stream);
            break;
        default: return RP_ERR_UNSUPPORTED;
        }
        if (rc != RP_OK) return rc;
    }
    return RP_OK;
}

}  // namespace

extern "C" {
int rp_scnet_forward(const rp_net_op* ops, int n_ops, void* stream) { return run_plan(ops, n_ops, stream); }
int rp_resnet18_8s_forward(const rp_net_op* ops, int n_ops, void* stream) { return run_plan(ops, n_ops, stream); }
}

"""Device-resident batched pipeline for BASELINE configs[2] / [3]: network forward -> matching primitives -> RPModule solve
for B scan pairs at once, and the batched form of RelativePoseEstimationViaCompletion (rpmodule.py:569-662).

The reference processes one pair at a time and crosses the host boundary four times per alternation step (SURVEY.md
section 3.1).  Here the hand-off (rpmodule.getMatchingPrimitive, :511-538) is one kernel (csrc/rp_keypoint.cu:
rp_gather_primitives) writing the rows rp_solve_batch reads; only the [B,4,4] poses come back to the host.  Keypoint pixel
locations are an input ([2B,K,2]; SIFT is OpenCV on the CPU and stays outside).

Image order everywhere: [source_0, target_0, source_1, target_1, ...] (pair b = images 2b, 2b+1), the order
SCNet.forward takes (evaluation.py:240).
"""
import ctypes

import numpy as np

from . import _lib, solver as _solver, util as _util


class _Stager(object):
    """Host arrays -> device tensors through ONE reusable pinned buffer: copy threads fill it in 8 MB pieces and every piece
    is sent (cudaMemcpyAsync) as soon as it is complete, so the host copy of piece k+1 overlaps the DMA of piece k.  A plain
    ``torch.as_tensor(a).to(dev)`` of pageable memory moves the 180 MB of a 32-pair batch at ~7 GB/s; this path is bounded by
    the PCIe link instead."""
    PIECE = 8 << 20

    def __init__(self):
        self.buf, self.event, self.pool = None, None, None

    def upload(self, arrays, dev, stream=None):
        """arrays: numpy arrays / torch tensors; returns device tensors of the same shapes and dtypes (CUDA tensors pass through).
        ``stream``: a side stream for the copies (the caller makes its compute stream wait on ``self.event``)."""
        import torch
        from concurrent.futures import ThreadPoolExecutor
        out, jobs, total = [None] * len(arrays), [], 0
        for i, a in enumerate(arrays):
            if torch.is_tensor(a):
                if a.is_cuda:
                    out[i] = a.to(dev)
                    continue
                a = a.detach().numpy()
            a = np.ascontiguousarray(a)
            jobs.append((i, a, total))
            total += (a.nbytes + 255) // 256 * 256
        if not jobs:
            return out
        if self.event is not None:
            self.event.synchronize()                       # the previous upload's copies have left the buffer
        if self.buf is None or self.buf.numel() < total:
            self.buf = torch.empty((int(total * 1.1),), dtype=torch.uint8).pin_memory()
        if self.pool is None:
            self.pool = ThreadPoolExecutor(max_workers=4)
        hostv = self.buf.numpy()
        futs = []
        for i, a, off in jobs:
            flat = a.reshape(-1).view(np.uint8)
            dst = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, device=dev)
            out[i] = dst
            dflat = dst.view(-1).view(torch.uint8)
            for lo in range(0, flat.size, self.PIECE):
                hi = min(flat.size, lo + self.PIECE)
                futs.append((self.pool.submit(np.copyto, hostv[off + lo:off + hi], flat[lo:hi]), dflat, off, lo, hi))
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream()
            st = stream or cur
            if stream is not None:
                stream.wait_stream(cur)                    # the destination tensors were allocated on the compute stream
            with torch.cuda.stream(st):
                for fut, dflat, off, lo, hi in futs:
                    fut.result()
                    dflat[lo:hi].copy_(self.buf[off + lo:off + hi], non_blocking=True)
                self.event = torch.cuda.Event()
                self.event.record(st)
        return out


_stagers = [_Stager(), _Stager()]      # two, so that filling one pinned buffer overlaps the DMA out of the other
_copy_streams = {}


def gather_primitives(feat, depth, normal, pts, weights, dataset, raw=False):
    """feat: CUDA float32 [2B,C,160,640] (may be a channel slice of the network output); depth [2B,160,640], normal
    [2B,160,640,3] (CUDA, converted to float64); pts [2B,K,2] pixel (x,y) float64; weights [2B,K] (1.0 observed / 0.99).
    -> DeviceBatch for PoseSolver.solve_device (``raw``: the gathered arrays themselves, [2B,K,.] in image order)."""
    import torch
    lib = _lib.load()
    n_img, C = feat.shape[0], feat.shape[1]
    assert feat.is_cuda and feat.dtype == torch.float32 and feat.shape[2] == 160 and feat.shape[3] == 640
    assert feat.stride(3) == 1 and feat.stride(2) == 640 and feat.stride(1) == 160 * 640, "feature maps must be contiguous planes"
    dev = feat.device
    depth = depth.to(dev).to(torch.float64).contiguous()
    normal = normal.to(dev).to(torch.float64).contiguous()
    pts = torch.as_tensor(pts, dtype=torch.float64).to(dev).contiguous()
    weights = torch.as_tensor(weights, dtype=torch.float64).to(dev).contiguous()
    K = pts.shape[1]
    B = n_img // 2
    pc = torch.empty((n_img, K, 3), dtype=torch.float64, device=dev)
    nn = torch.empty((n_img, K, 3), dtype=torch.float64, device=dev)
    desc = torch.empty((n_img, K, C), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.rp_gather_primitives(feat.data_ptr(), C, feat.stride(0), depth.data_ptr(), normal.data_ptr(), pts.data_ptr(),
                                            n_img, K, _util.dataset_id(dataset), pc.data_ptr(), nn.data_ptr(), desc.data_ptr(),
                                            torch.cuda.current_stream().cuda_stream), "rp_gather_primitives")
    if raw:
        return pc, nn, desc, weights

    return _batch_from_gathered(pc, nn, desc, weights)


def _batch_from_gathered(pc, nn, desc, weights):
    """[2B,K,3] positions / normals, [2B,K,C] descriptors, [2B,K] weights in image order -> DeviceBatch of B pairs."""
    n_img, K = pc.shape[0], pc.shape[1]
    B = n_img // 2

    def side(a, k):
        return a.view(B, 2, K, -1)[:, k].reshape(B * K, -1)
    return _solver.DeviceBatch.from_device_arrays(side(pc, 0), side(nn, 0), side(desc, 0), weights.view(B, 2, K)[:, 0].reshape(-1),
                                                  side(pc, 1), side(nn, 1), side(desc, 1), weights.view(B, 2, K)[:, 1].reshape(-1), B)


def _net_forward(net, inp, heads=None):
    """SCNet forward without the defensive output copy when ``net`` is this package's SCNet: the engine's output buffer is
    consumed (blend, gather) before the next forward, and only the output heads the caller reads are computed (``heads``).  Any
    other module is simply called."""
    from .model.mymodel import SCNet
    if isinstance(net, SCNet):
        from . import scnet_engine
        if net._engine is None:
            net._engine = scnet_engine.ScnetEngine(net)
        return net._engine.forward(inp, borrow=True, heads=heads)
    return net(inp)


def solve_from_maps(feat, depth, normal, pts, weights, para, dataset, solver=None):
    """Poses [B,4,4] float64 (numpy) for B pairs from descriptor / depth / normal maps and keypoints; one gather kernel +
    one fused solver launch."""
    solver = solver or _solver.default_solver(feat.device)
    d = gather_primitives(feat, depth, normal, pts, weights, dataset)
    T, status, _ = solver.solve_device_checked(d, [_solver.params_from_opts(para)])
    return T.cpu().numpy()


def RelativePoseEstimationViaCompletion_batch(net, rgb, norm, depth, pts, weights, args, chunk=None):
    """Batched rpmodule.RelativePoseEstimationViaCompletion (rpmodule.py:569-662) for B pairs with given keypoints.

    rgb [2B,160,640,3], norm [2B,160,640,3], depth [2B,160,640] (numpy or tensors; the complete scans), pts [2B,K,2],
    weights [2B,K]; args as in the reference (snumclass, featureDim, outputType, maskMethod, alterStep, dataset, para with
    per-step sigma arrays).  Per alternation: a batched warp of the 2B views, the SCNet forward, the blend and the gather run
    over ``chunk`` pairs at a time (default: all B; the network's activation buffers are sized by it), then ONE solve over all B
    pairs -- the solver's cost per pair falls with the batch (a pair alone occupies one SM).  Scans are uploaded once, chunk by
    chunk on a copy stream, so the upload of chunk c+1 overlaps the first network pass of chunk c.  Returns [B,4,4] float64."""
    import copy
    import torch
    dev = next(net.parameters()).device
    idx_f = 0
    for key, n in (('rgb', 3), ('n', 3), ('d', 1), ('s', args.snumclass)):
        if key in args.outputType:
            idx_f += n
    n_img = len(rgb)
    B = n_img // 2
    chunk = B if not chunk else max(1, min(int(chunk), B))
    bounds = [(c0, min(B, c0 + chunk)) for c0 in range(0, B, chunk)]
    as64 = lambda a: a if torch.is_tensor(a) else np.asarray(a, dtype=np.float64)
    pts, weights = as64(pts), as64(weights)
    solver = _solver.default_solver(dev)
    state = {}

    def prepare(ci, lo, hi):
        """Upload the scans of pairs lo..hi and build their static network input (own view in channels 0..7)."""
        sl = slice(2 * lo, 2 * hi)
        with torch.cuda.device(dev):
            cs = _copy_streams.setdefault(str(dev), torch.cuda.Stream(device=dev)) if len(bounds) > 1 else None
            stg = _stagers[ci % 2]
            rgb_d, norm_gt, depth_gt, pts_d, w_d = stg.upload([rgb[sl], norm[sl], depth[sl], pts[sl], weights[sl]], dev, stream=cs)
            if cs is not None and stg.event is not None:
                torch.cuda.current_stream().wait_event(stg.event)
        full = torch.cat((rgb_d.float(), norm_gt.float(), depth_gt.float().unsqueeze(3)), 3).permute(0, 3, 1, 2).contiguous()   # [2b,7,h,w]
        vw, m, _geow = _util.apply_mask(full, args.maskMethod)
        inp = torch.empty((2 * (hi - lo), 16, 160, 640), dtype=torch.float32, device=dev)
        inp[:, :7] = vw                                               # network input: own view | partner warped into this frame
        inp[:, 7] = (vw[:, 6] != 0).float()
        swap = (torch.arange(2 * (hi - lo), device=dev) ^ 1).to(torch.int32)   # the other scan of the pair
        return dict(inp=inp, mask=m[:, 0].contiguous(), norm_gt=norm_gt, depth_gt=depth_gt, pts=pts_d.double(), w=w_d.double(), swap=swap)

    with torch.no_grad():
        R_hat = np.tile(np.eye(4), (B, 1, 1))
        for alter_ in range(args.alterStep):
            para_this = copy.copy(args.para)
            for name in ('sigmaAngle1', 'sigmaAngle2', 'sigmaDist', 'sigmaFeat'):
                setattr(para_this, name, getattr(args.para, name)[alter_])
            parts = []
            for ci, (lo, hi) in enumerate(bounds):
                if alter_ == 0:
                    state[ci] = prepare(ci, lo, hi)
                S = state[ci]
                # view i receives the other scan warped into its frame: sources get the target moved by inv(R), targets the
                # source moved by R (rpmodule.py:616-617); identity -> zeros inside the kernel
                Rs = np.empty((2 * (hi - lo), 4, 4))
                Rs[0::2] = np.linalg.inv(R_hat[lo:hi])
                Rs[1::2] = R_hat[lo:hi]
                _util.warping_device(S['inp'], Rs, args.dataset, out=S['inp'][:, 8:], src_index=S['swap'])   # reads channels 0..7, writes 8..15
                f = _net_forward(net, S['inp'], heads=('n', 'd', 'f'))    # :619-623; the alternation reads f[3:7] (:628-634) and the descriptors
                nrm2, dep2 = _util.blend_completion_device(f, S['mask'], S['norm_gt'], S['depth_gt'])          # :628-634
                parts.append(gather_primitives(f[:, idx_f:idx_f + args.featureDim], dep2, nrm2, S['pts'], S['w'], args.dataset, raw=True))
            d = _batch_from_gathered(*[torch.cat([p[k] for p in parts], 0) if len(parts) > 1 else parts[0][k] for k in range(4)])
            T, status, _ = solver.solve_device_checked(d, [_solver.params_from_opts(para_this)])
            R_hat = T.cpu().numpy()
    return R_hat

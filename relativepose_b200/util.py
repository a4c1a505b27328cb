"""Drop-in for the view-geometry helpers of the reference's util.py that sit on both sides of the completion network in
every alternation of RelativePoseEstimationViaCompletion, backed by csrc/rp_warp.cu:

  warping(view, R, dataList)            util.py:94-172   (+ reproj_helper :537-749, depth2pc :468-523)
  Pano2PointCloud(depth, dataList)      util.py:751-811
  apply_mask(x, maskMethod, *arg)       util.py:209-232
  blend_completion(...)                 RPModule/rpmodule.py:628-634 (not a function of its own in the reference)

The ``*_device`` forms take / return CUDA tensors and whole batches (one launch pair per call, no host round trip);
the reference-named functions keep the reference's numpy signatures on top of them.  No CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib

_DATASETS = ('suncg', 'matterport', 'scannet')


def dataset_id(dataList):
    for i, k in enumerate(_DATASETS):
        if k in dataList:
            return i
    raise ValueError("unknown dataset %r" % (dataList,))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("relativepose_b200.util needs a CUDA device (no CPU fallback)")
    return torch


_ws = {}


def warping_device(views, R, dataList, out=None, src_index=None):
    """views: CUDA float32 [B,8,160,640] (or a [B,>=8,160,640] tensor whose first 8 channels are the view); R: [B,4,4]
    (numpy / tensor, float64) -> CUDA float32 [B,8,160,640].  ``src_index`` (int32 CUDA [B]): output b is views[src_index[b]]
    warped by R[b].  ``out``: a [B,8,160,640] tensor or a channel slice of a wider [B,C,160,640] one (image stride C*H*W)."""
    torch = _torch()
    lib = _lib.load()
    if views.dim() != 4 or views.shape[1] < 8 or views.shape[2] != 160 or views.shape[3] != 640:
        raise ValueError("expected views [B,8,160,640]")
    if views.dtype != torch.float32 or views.stride(3) != 1 or views.stride(2) != 640 or views.stride(1) != 160 * 640:
        views = views[:, :8].contiguous().float()
    B = views.shape[0]
    dev = views.device
    Rt = torch.as_tensor(np.ascontiguousarray(np.asarray(R.cpu() if hasattr(R, 'cpu') else R, dtype=np.float64).reshape(B, 16))).to(dev)
    if out is None:
        out = torch.empty((B, 8, 160, 640), dtype=torch.float32, device=dev)
    assert out.dtype == torch.float32 and out.shape[0] == B and out.shape[1] == 8 and out.stride(3) == 1 and out.stride(2) == 640 \
        and out.stride(1) == 160 * 640, "out must be [B,8,160,640] planes (possibly a channel slice of a wider tensor)"
    need = ctypes.c_size_t(0)
    _lib.check(lib.rp_warp_workspace_bytes(B, ctypes.byref(need)), "rp_warp_workspace_bytes")
    key = (str(dev), 'warp')
    ws = _ws.get(key)
    if ws is None or ws.numel() < need.value:
        ws = torch.empty((max(need.value, 1),), dtype=torch.uint8, device=dev)
        _ws[key] = ws
    with torch.cuda.device(dev):
        _lib.check(lib.rp_warp_views_ex(views.data_ptr(), views.stride(0), src_index.data_ptr() if src_index is not None else None,
                                        Rt.data_ptr(), B, dataset_id(dataList), out.data_ptr(), out.stride(0), ws.data_ptr(),
                                        ws.numel(), torch.cuda.current_stream().cuda_stream), "rp_warp_views_ex")
    return out


def warping(view, R, dataList):
    """util.warping (util.py:94-172): view [1,8,160,640] numpy (rgb, normal, depth, valid), R [4,4] -> [1,8,160,640].
    Returns float64 numpy like the reference; the values are the reference's float64 results rounded to float32 (what
    its only caller keeps, ``torch_op.v(util.warping(...))``, rpmodule.py:616-617).  Identity R -> zeros (util.py:95-96)."""
    torch = _torch()
    v = torch.as_tensor(np.ascontiguousarray(view, dtype=np.float32)).cuda()
    return warping_device(v, np.asarray(R, dtype=np.float64)[None], dataList).cpu().numpy().astype(np.float64)


def pano2pointcloud_device(depth, dataList):
    """depth: CUDA float32 [B,160,640] -> (pc [B,3,102400] float64, valid [B,102400] uint8), reference point order."""
    torch = _torch()
    lib = _lib.load()
    depth = depth.contiguous().float()
    B = depth.shape[0]
    pc = torch.empty((B, 3, 160 * 640), dtype=torch.float64, device=depth.device)
    valid = torch.empty((B, 160 * 640), dtype=torch.uint8, device=depth.device)
    with torch.cuda.device(depth.device):
        _lib.check(lib.rp_pano2pc(depth.data_ptr(), B, dataset_id(dataList), pc.data_ptr(), valid.data_ptr(),
                                  torch.cuda.current_stream().cuda_stream), "rp_pano2pc")
    return pc, valid


def Pano2PointCloud(depth, dataList):
    """util.Pano2PointCloud (util.py:751-811): depth [160,640] numpy -> [3,n] float64 (scannet: depth != 0 only)."""
    torch = _torch()
    assert depth.shape[0] == 160 and depth.shape[1] == 640
    pc, valid = pano2pointcloud_device(torch.as_tensor(np.ascontiguousarray(depth, dtype=np.float32)).cuda()[None], dataList)
    pc = pc[0]
    if 'scannet' in dataList:
        pc = pc[:, valid[0].bool()]               # compaction of the kept points (stable: reference order)
    return pc.cpu().numpy()


def apply_mask(x, maskMethod, *arg):
    """util.apply_mask (util.py:209-232).  x: torch [n,c,h,w].  Returns the reference's 3-tuple (masked x, mask [n,1,h,w],
    geow [n,1,h,w]): 'second' observes skybox face 1 (columns h..2h) and geow = exp(-d/(2*0.7^2)) of the normalised column
    distance d to the nearest face border, zero on the observed face (:215-222); 'kinect' observes a 66x88 window of that
    face and geow = 1 - mask (:223-229).  Any other maskMethod leaves mask and geow all zero, as the reference does."""
    import torch
    h, w = x.shape[2], x.shape[3]
    m = torch.zeros((x.shape[0], 1, h, w), dtype=x.dtype, device=x.device)
    geow = torch.zeros((x.shape[0], 1, h, w), dtype=x.dtype, device=x.device)
    if maskMethod == 'second':
        m[:, :, :h, h:2 * h] = 1
        xs = np.arange(w)
        dist = np.stack((np.abs(xs - h), np.abs(xs - 2 * h), np.abs(xs - w - h), np.abs(xs - w - 2 * h)), 0).min(0) / h
        dist = np.exp(-dist / (2 * 0.7 ** 2))
        dist[h:2 * h] = 0
        geow[:] = torch.as_tensor(dist, dtype=x.dtype, device=x.device)[None, None, None, :]
    elif maskMethod == 'kinect':
        assert w == 640 and h == 160
        dw, dh = int(89.67 // 2), int(67.25 // 2)
        m[:, :, 80 - dh:80 + dh, 160 + 80 - dw:160 + 80 + dw] = 1
        geow = 1 - m
    return x * m, m, geow


def blend_completion_device(f, mask, norm_gt, depth_gt):
    """rpmodule.py:628-634 on the device.  f: CUDA float32 [B,C,160,640]; mask: [B,160,640] float32; norm_gt [B,160,640,3],
    depth_gt [B,160,640] float64 or float32 (the result has their type) -> (normal [B,160,640,3], depth [B,160,640])."""
    torch = _torch()
    lib = _lib.load()
    f = f.contiguous().float()
    B, C = f.shape[0], f.shape[1]
    dt = torch.float64 if norm_gt.dtype == torch.float64 or depth_gt.dtype == torch.float64 else torch.float32
    mask = mask.to(f.device).contiguous().float()
    norm_gt = norm_gt.to(f.device).to(dt).contiguous()
    depth_gt = depth_gt.to(f.device).to(dt).contiguous()
    nrm, dep = torch.empty_like(norm_gt), torch.empty_like(depth_gt)
    with torch.cuda.device(f.device):
        _lib.check(lib.rp_blend_completion(f.data_ptr(), C, mask.data_ptr(), norm_gt.data_ptr(), depth_gt.data_ptr(),
                                           int(dt == torch.float64), B, nrm.data_ptr(), dep.data_ptr(),
                                           torch.cuda.current_stream().cuda_stream), "rp_blend_completion")
    return nrm, dep

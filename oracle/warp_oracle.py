"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's view warping.

Follows /root/reference/util.py: ``warping`` (:94-172), ``reproj_helper`` (:537-749), ``Pano2PointCloud`` (:751-811),
``depth2pc`` (:468-523) and the blend / re-normalise step of ``RelativePoseEstimationViaCompletion``
(RPModule/rpmodule.py:628-634).  Only tests/, __graft_entry__.smoke() and bench.py's CPU arm may import this module;
the product path (relativepose_b200/util.py -> csrc/rp_warp.cu) never does.

Pinned against the reference itself: tests/golden/make_warp_golden.py runs the unmodified util.warping on seeded views
and commits masks / sums / sampled values (tests/golden/warp_golden.npz); tests/test_warp_oracle.py checks this file
against them and, when /root/reference is present, against the live reference bit for bit.

The restatement is organised differently from the reference (one face table instead of three copies of the same
~70 lines per dataset) but performs the same float64 operations in the same order on the same numpy calls.
"""
import numpy as np

H = 160

# Rs of util.py:539-543 (identical in every branch): camera-to-world rotations of the four skybox faces
_RS = np.zeros([4, 3, 3])
_RS[0] = np.eye(3)
_RS[1] = np.array([[0, 0, -1], [0, 1, 0], [1, 0, 0]])
_RS[2] = np.array([[-1, 0, 0], [0, 1, 0], [0, 0, -1]])
_RS[3] = np.array([[0, 0, 1], [0, 1, 0], [-1, 0, 0]])

# which Rs the panorama column block `slot` (0..3 = columns slot*160..) looks through:
#   suncg (util.py:546-602): slots use Rs[0], Rs[1], Rs[2], Rs[3]; matterport / scannet (:615-672, :686-743): Rs[3], Rs[0], Rs[1], Rs[2]
FACE_OF_SLOT = {'suncg': (0, 1, 2, 3), 'matterport': (3, 0, 1, 2), 'scannet': (3, 0, 1, 2)}


def dataset_key(dataList):
    for k in ('suncg', 'matterport', 'scannet'):
        if k in dataList:
            return k
    raise ValueError("unknown dataset %r" % (dataList,))


def pano2pointcloud(depth, dataList):
    """util.py:751-811.  depth [160,640] -> [3, n] float64 (n = 102400; scannet: only depth != 0, per face)."""
    key = dataset_key(dataList)
    assert depth.shape[0] == 160 and depth.shape[1] == 640
    w, h = depth.shape[1] // 4, depth.shape[0]
    ys, xs = np.meshgrid(range(h), range(w), indexing='ij')
    ys, xs = (0.5 - ys / h) * 2, (xs / w - 0.5) * 2
    pc = []
    for i in range(4):
        zs = depth[:, i * w:(i + 1) * w].flatten()
        if key == 'scannet':
            mask = (zs != 0)
            zs = zs[mask]
            ys_this, xs_this = ys.flatten()[mask] * zs / (1.1895 * 2), xs.flatten()[mask] * zs / (0.8921875 * 2)
        else:
            ys_this, xs_this = ys.flatten() * zs, xs.flatten() * zs
        pc_this = np.concatenate((xs_this, ys_this, -zs)).reshape(3, -1)
        R = _RS[i] if key == 'suncg' else _RS[(i - 1) % 4]
        pc.append(np.matmul(R, pc_this))
    return np.concatenate(pc, 1)


def depth2pc(depth, dataList):
    """util.py:468-523 for the shapes the warp uses.  -> (pc [n,3] float64, mask [H*W] bool)."""
    key = dataset_key(dataList)
    h, w = depth.shape
    ys, xs = np.meshgrid(range(h), range(w), indexing='ij')
    if key in ('suncg', 'matterport'):
        assert w == 160 and h == 160
        ys, xs = (0.5 - ys / h) * 2, (xs / h - 0.5) * 2
    else:
        assert (h, w) in ((66, 88), (480, 640))
        ys, xs = (0.5 - ys / h) * 2, (xs / w - 0.5) * 2
    zs = depth.flatten()
    mask = (zs != 0)
    zs = zs[mask]
    if key == 'scannet' and (h, w) == (480, 640):
        xs = xs.flatten()[mask] * zs / (0.8921875 * 2)
        ys = ys.flatten()[mask] * zs / (1.1895 * 2)
        pc = np.stack((xs, ys, -zs), 1)
    else:
        xs = xs.flatten()[mask] * zs
        ys = ys.flatten()[mask] * zs
        if key == 'scannet':
            pc = np.stack((xs * w / 160, ys * h / 160, -zs), 1)
        else:
            pc = np.stack((xs, ys, -zs), 1)
    if key == 'suncg':
        pc = np.matmul(_RS[1], pc.T).T          # "assume second view" (util.py:483-484)
    return pc, mask


def reproj(pct, values, out_shape, mode, dataList):
    """util.py:537-749.  pct [3,n] float64 points in the target frame; values [n,3] (mode 'color'/'normal') or ignored
    (mode 'depth': the value is the depth along the face axis).  Scatter with numpy's last-write-wins semantics, faces
    written in the reference's order front, left, back, right."""
    key = dataset_key(dataList)
    h = out_shape[0]
    per_slot = []
    for slot in range(4):
        R = _RS[FACE_OF_SLOT[key][slot]]
        tp = pct.copy() if (key == 'suncg' and slot == 0) else np.matmul(R.T, pct)
        tp[:2, :] /= (np.abs(tp[2, :]) + 1e-32)
        inter = (tp[2, :] < 0) * (np.abs(tp[0, :]) < 1) * (np.abs(tp[1, :]) < 1)
        val = values[inter, :] if mode in ('color', 'normal') else -tp[2, inter]
        coord = tp[:2, inter]
        coord[0, :] = (coord[0, :] + 1) * 0.5 * h
        coord[1, :] = (1 - coord[1, :]) * 0.5 * h
        coord = coord.round().clip(0, h - 1).astype('int')
        coord[0, :] += h * slot
        per_slot.append((coord, val))
    proj = np.zeros(out_shape)
    for slot in (0, 3, 2, 1):                   # f, l, b, r (util.py:604-608)
        coord, val = per_slot[slot]
        proj[coord[1, :], coord[0, :]] = val
    return proj


def warping(view, R, dataList):
    """util.py:94-172.  view [1,8,160,640] (rgb, normal, depth, mask), R [4,4] -> [1,8,160,640] float64; identity R ->
    zeros (the reference returns a float32 CUDA tensor of zeros there, :95-96)."""
    key = dataset_key(dataList)
    if np.linalg.norm(R - np.eye(4)) == 0:
        return np.zeros(view.shape)
    h = 160
    rgb = view[0, 0:3, :, :].transpose(1, 2, 0)
    normal = view[0, 3:6, :, :].transpose(1, 2, 0)
    depth = view[0, 6, :, :]
    if key == 'suncg':
        pct = pano2pointcloud(depth, 'suncg')
        colorpct = rgb[:, h:2 * h, :].reshape(-1, 3)
        normalpct = normal[:, h:2 * h, :].reshape(-1, 3)
        pct_reproj = np.matmul(R, np.concatenate((pct, np.ones([1, pct.shape[1]]))))[:3, :]
        pct_reproj = pct_reproj[:, h * h:h * h * 2]
    else:
        if key == 'matterport':
            win = (slice(0, h), slice(h, 2 * h))
        else:
            assert view.shape[2] == 160 and view.shape[3] == 640
            win = (slice(80 - 33, 80 + 33), slice(160 + 80 - 44, 160 + 80 + 44))
        pct, mask = depth2pc(depth[win], key)
        colorpct = rgb[win].reshape(-1, 3)[mask, :]
        normalpct = normal[win].reshape(-1, 3)[mask, :]
        pct_reproj = np.matmul(R, np.concatenate((pct.T, np.ones([1, pct.shape[0]]))))[:3, :]
    normalpct = np.matmul(R[:3, :3], normalpct.T).T
    s2t_rgb = reproj(pct_reproj, colorpct, rgb.shape, 'color', key)
    s2t_n = reproj(pct_reproj, normalpct, rgb.shape, 'normal', key)
    s2t_d = reproj(pct_reproj, None, rgb.shape[:2], 'depth', key)
    s2t_mask = (s2t_d != 0).astype('int')
    out = np.concatenate((s2t_rgb, s2t_n, np.expand_dims(s2t_d, 2), np.expand_dims(s2t_mask, 2)), 2)
    return np.expand_dims(out, 0).transpose(0, 3, 1, 2)


def blend_completion(f, mask, norm_gt, depth_gt, eps=1e-12):
    """RPModule/rpmodule.py:628-634: keep the observed region, fill the rest with the network's prediction.
    f [C,160,640] float32 (channels 3:6 normal, 6 depth), mask [160,640,1], norm_gt [160,640,3], depth_gt [160,640]."""
    nrm = (1 - mask) * f[3:6].transpose(1, 2, 0) + mask * norm_gt
    nrm = nrm / (np.linalg.norm(nrm, axis=2, keepdims=True) + eps)
    dep = (1 - mask[:, :, 0]) * f[6] + mask[:, :, 0] * depth_gt
    return nrm, dep

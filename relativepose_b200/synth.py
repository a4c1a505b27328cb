"""Seeded synthetic matching primitives and panoramas.

The reference ships no data (datasets/checkpoints are download-only,
README.md:24-28), so the benchmark, the oracle and the parity tests all draw
from this generator.  A *record* uses the wire format of the reference's
primitive cache (trainRelativePoseModuleRecFD.py:207-208).
"""
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

FEAT_DIM = 32  # descriptor width (model/mymodel.py:228, rpmodule.py:323)


def shipped_params(dataset="suncg"):
    """Rows [sigmaAngle1, sigmaAngle2, sigmaDist, sigmaFeat] of the parameter file
    the reference ships (data/relativePoseModule/final_param_<ds>_rlevel_3.txt,
    read by evaluation.py:95-101)."""
    path = os.path.join(_DATA, "final_param_%s_rlevel_3.txt" % dataset)
    return np.loadtxt(path).reshape(-1, 4)


def random_rigid(rs):
    axis = rs.randn(3)
    axis /= np.linalg.norm(axis)
    ang = rs.uniform(0.3, 2.5)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)
    t = rs.randn(3) * 0.5
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def _unit(v):
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def make_pair(seed, n_s, n_t=None, inlier_frac=0.5, pos_noise=0.005, feat_noise=0.02):
    """One scan pair's matching primitives.

    source points U(-3,3)^3 m, unit normals, descriptors tanh(N(0,1)) float32
    (SCNet's f-head is tanh bounded, model/mymodel.py:374-375); the first
    floor(inlier_frac*min(n_s,n_t)) target rows are the rigidly moved source rows
    plus noise, the rest are independent draws; weights are 1.0 (observed) w.p.
    0.5, else 0.99 (rputil.py:229-235).
    """
    if n_t is None:
        n_t = n_s
    rs = np.random.RandomState(seed)
    T = random_rigid(rs)
    R, t = T[:3, :3], T[:3, 3]
    pc_s = rs.uniform(-3, 3, size=(n_s, 3))
    nrm_s = _unit(rs.randn(n_s, 3))
    feat_s = np.tanh(rs.randn(n_s, FEAT_DIM))
    pc_t = rs.uniform(-3, 3, size=(n_t, 3))
    nrm_t = _unit(rs.randn(n_t, 3))
    feat_t = np.tanh(rs.randn(n_t, FEAT_DIM))
    n_in = int(np.floor(inlier_frac * min(n_s, n_t)))
    if n_in > 0:
        pc_t[:n_in] = pc_s[:n_in] @ R.T + t + rs.randn(n_in, 3) * pos_noise
        nrm_t[:n_in] = nrm_s[:n_in] @ R.T
        feat_t[:n_in] = feat_s[:n_in] + rs.randn(n_in, FEAT_DIM) * feat_noise
    w_s = np.where(rs.rand(n_s) < 0.5, 1.0, 0.99)
    w_t = np.where(rs.rand(n_t) < 0.5, 1.0, 0.99)
    return {
        "pc_src": pc_s, "normal_src": nrm_s, "feat_src": feat_s.astype(np.float32), "weight_src": w_s,
        "pc_tgt": pc_t, "normal_tgt": nrm_t, "feat_tgt": feat_t.astype(np.float32), "weight_tgt": w_t,
        "R_gt": T,
    }


def record_to_dicts(rec):
    """(dataS, dataT) in the layout RelativePoseEstimation_helper takes (rpmodule.py:317-325)."""
    s = {"pc": rec["pc_src"], "normal": rec["normal_src"], "feat": rec["feat_src"], "weight": rec["weight_src"]}
    t = {"pc": rec["pc_tgt"], "normal": rec["normal_tgt"], "feat": rec["feat_tgt"], "weight": rec["weight_tgt"]}
    return s, t


def keypoints_for_nominal_N(N, topk=5):
    """n_s = n_t = ceil(N/topK): nominal 64..2048 -> N_actual 65,130,260,515,1025,2050."""
    return -(-N // topk)


def make_batch(first_seed, count, n_s, n_t=None, **kw):
    return [make_pair(first_seed + i, n_s, n_t, **kw) for i in range(count)]


def make_panorama_pair(seed, dataset="suncg", H=160, W=640):
    """Synthetic RGB-D skybox pair assembled like evaluation.py:217-239: per scan
    [rgb U(0,1), unit normals, depth U(0.5,5) with 5% invalid zeros] masked to the observed face(s)
    (util.py:209-232: 'second' = columns h..2h for suncg/matterport, 'kinect' = 66x88 window for scannet),
    plus the validity mask channel; the warped-other-view half is zero on the first alternation step
    (util.py:95-96).  Returns float32 [2,16,H,W]."""
    rs = np.random.RandomState(seed)
    out = np.zeros([2, 16, H, W], dtype=np.float32)
    for v in range(2):
        rgb = rs.uniform(0, 1, size=(3, H, W))
        nrm = rs.randn(3, H, W)
        nrm /= np.linalg.norm(nrm, axis=0, keepdims=True)
        depth = rs.uniform(0.5, 5.0, size=(1, H, W))
        depth[rs.rand(1, H, W) < 0.05] = 0.0
        full = np.concatenate((rgb, nrm, depth), 0)
        mask = np.zeros([1, H, W])
        if "scannet" in dataset:
            dw, dh = int(89.67 // 2), int(67.25 // 2)
            mask[:, 80 - dh:80 + dh, 160 + 80 - dw:160 + 80 + dw] = 1
        else:
            mask[:, :H, H:2 * H] = 1
        view = full * mask
        valid = (view[6:7] != 0).astype(np.float64)          # evaluation.py:225-228 / rpmodule.py:609-612
        out[v, 0:7] = view
        out[v, 7:8] = valid
    return out


def make_warp_view(seed, dataset="suncg"):
    """One observed view [1,8,160,640] float32 as RelativePoseEstimationViaCompletion assembles it (rpmodule.py:599-612):
    rgb, unit normals, a smooth depth field with 5 % invalid zeros, masked to the observed face / Kinect window, plus the
    validity channel.  The depth is smooth (not white noise) so that the warp produces coherent surfaces with many
    many-to-one collisions -- the case the last-write-wins scatter of util.reproj_helper (util.py:603-608) is about."""
    rs = np.random.RandomState(seed)
    view = np.zeros((1, 8, 160, 640), np.float32)
    view[0, 0:3] = rs.rand(3, 160, 640)
    n = rs.randn(3, 160, 640)
    view[0, 3:6] = n / np.linalg.norm(n, axis=0)
    yy, xx = np.mgrid[0:160, 0:640]
    d = 2.5 + 1.5 * np.sin(xx / 37.0 + seed) * np.cos(yy / 23.0) + 0.2 * rs.rand(160, 640)
    d[rs.rand(160, 640) < 0.05] = 0
    view[0, 6] = d
    m = np.zeros((160, 640), np.float32)
    if "scannet" in dataset:
        m[80 - 33:80 + 33, 160 + 80 - 44:160 + 80 + 44] = 1
    else:
        m[:, 160:320] = 1
    view[0, :7] *= m
    view[0, 7] = (view[0, 6] != 0)
    return view


def make_pose(seed, max_angle=2.0, t_sigma=0.5):
    """Seeded rigid transform [4,4] float64 (Rodrigues rotation about a random axis, angle U(0.2, max_angle))."""
    rs = np.random.RandomState(100 + seed)
    ax = rs.randn(3)
    ax /= np.linalg.norm(ax)
    ang = rs.uniform(0.2, max_angle)
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    R = np.eye(4)
    R[:3, :3] = np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * K @ K
    R[:3, 3] = rs.randn(3) * t_sigma
    return R


def make_feature_map(seed, C=32, H=160, W=640):
    """Smooth, tanh-bounded synthetic descriptor map [C,H,W] float32 (the f-head of SCNet is tanh bounded and spatially
    smooth): bilinearly zoomed low-resolution noise plus a little per-pixel noise.  numpy/scipy only, deterministic."""
    from scipy import ndimage
    rs = np.random.RandomState(seed)
    lo = rs.randn(C, H // 8, W // 8)
    f = ndimage.zoom(lo, (1, 8, 8), order=1)
    return np.tanh(f + 0.05 * rs.randn(C, H, W)).astype(np.float32)


def make_texture_image(seed, H=160, W=640):
    """uint8 [H,W,3] image of random soft blobs (something a SIFT detector finds keypoints on)."""
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    img = np.zeros((H, W, 3))
    for _ in range(int(150 * H * W / (160 * 640))):
        x, y, r = rs.randint(0, W), rs.randint(0, H), rs.randint(3, 14)
        m = 1.0 / (1.0 + np.exp((np.sqrt((xx - x) ** 2 + (yy - y) ** 2) - r) / 1.2))
        img = img * (1 - m[:, :, None]) + m[:, :, None] * rs.rand(3)
    return (img * 255).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------------
# A geometrically consistent scan pair: the inside of a textured box room rendered into the reference's 160x640 four-face
# skybox (util.Pano2PointCloud 'suncg' convention, util.py:755-773: face i looks along -z of Rs[i], x = (col/h-0.5)*2,
# y = (0.5-row/h)*2, depth = distance along the face axis).  Used by the RelativePoseEstimationViaCompletion golden
# (tests/golden/make_via_completion_golden.py) and the alternation benchmark: the two scans see the same walls, so the
# warp, the blend and the matcher get consistent geometry (random panoramas do not overlap in 3-D).
_SKYBOX_RS = np.array([[[1, 0, 0], [0, 1, 0], [0, 0, 1]],
                       [[0, 0, -1], [0, 1, 0], [1, 0, 0]],
                       [[-1, 0, 0], [0, 1, 0], [0, 0, -1]],
                       [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]], dtype=np.float64)


def _room_texture(rs, res=48):
    """Per-wall low-resolution colour tiles (piecewise constant: corners for a SIFT detector) plus smooth shading."""
    return rs.rand(6, res, res, 3), rs.uniform(0.5, 3.0, size=(6, 3, 2)), rs.uniform(0, 6.28, size=(6, 3, 2))


def render_room_scan(T_w2c, half, tex, H=160):
    """Render one scan.  T_w2c [4,4] world -> camera, half = box half-extents (x, y, z).  Returns the reference's scan dict
    {'rgb' [H,4H,3] in [0,1], 'norm' [H,4H,3] unit normals in the camera frame, 'depth' [H,4H]} (float64)."""
    tiles, freq, phase = tex
    res = tiles.shape[1]
    R_c2w = T_w2c[:3, :3].T
    o = -R_c2w @ T_w2c[:3, 3]                                   # camera centre in the world
    ys, xs = np.meshgrid(np.arange(H), np.arange(H), indexing='ij')
    ys, xs = (0.5 - ys / H) * 2, (xs / H - 0.5) * 2
    rgb = np.zeros((H, 4 * H, 3)); nrm = np.zeros((H, 4 * H, 3)); dep = np.zeros((H, 4 * H))
    for i in range(4):
        d_face = np.stack((xs.ravel(), ys.ravel(), -np.ones(H * H)), 0)          # z-depth 1 along the face axis
        d_w = R_c2w @ (_SKYBOX_RS[i] @ d_face)                                    # [3, H*H]
        with np.errstate(divide='ignore', invalid='ignore'):
            t_hi = (np.asarray(half)[:, None] - o[:, None]) / d_w
            t_lo = (-np.asarray(half)[:, None] - o[:, None]) / d_w
        t_ax = np.where(d_w > 0, t_hi, np.where(d_w < 0, t_lo, np.inf))           # exit distance per axis
        ax = np.argmin(t_ax, 0)
        t = t_ax[ax, np.arange(H * H)]
        p = o[:, None] + d_w * t                                                  # world hit point
        sign = np.sign(d_w[ax, np.arange(H * H)])
        wall = 2 * ax + (sign > 0)
        n_w = np.zeros((3, H * H)); n_w[ax, np.arange(H * H)] = -sign             # inward normal
        ua, va = (ax + 1) % 3, (ax + 2) % 3
        u = p[ua, np.arange(H * H)]; v = p[va, np.arange(H * H)]
        hu = np.asarray(half)[ua]; hv = np.asarray(half)[va]
        iu = np.clip(((u / hu * 0.5 + 0.5) * res).astype(int), 0, res - 1)
        iv = np.clip(((v / hv * 0.5 + 0.5) * res).astype(int), 0, res - 1)
        col = tiles[wall, iu, iv]                                                 # [H*H,3]
        shade = 0.5 + 0.5 * np.sin(freq[wall, :, 0] * u[:, None] + phase[wall, :, 0]) * np.sin(freq[wall, :, 1] * v[:, None] + phase[wall, :, 1])
        col = 0.7 * col + 0.3 * shade
        rgb[:, i * H:(i + 1) * H] = col.reshape(H, H, 3)
        nrm[:, i * H:(i + 1) * H] = (T_w2c[:3, :3] @ n_w).T.reshape(H, H, 3)
        dep[:, i * H:(i + 1) * H] = t.reshape(H, H)
    return {'rgb': rgb, 'norm': nrm, 'depth': dep}


def make_room_scan_pair(seed, max_yaw=0.6, max_shift=0.4, tex_res=48):
    """Two scans of one room.  Returns (data_s, data_t, R_gt) with R_gt [4,4] mapping source-camera to target-camera
    coordinates (the convention of rpmodule.py:616-617: ``warping(view_s, R_hat)``)."""
    rs = np.random.RandomState(seed)
    half = np.array([rs.uniform(2.0, 3.5), rs.uniform(1.2, 1.6), rs.uniform(2.0, 3.5)])
    tex = _room_texture(rs, tex_res)

    def cam():
        yaw, tilt = rs.uniform(-max_yaw, max_yaw), rs.uniform(-0.05, 0.05)
        cy, sy, ct, st = np.cos(yaw), np.sin(yaw), np.cos(tilt), np.sin(tilt)
        R = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, ct, -st], [0, st, ct]])
        T = np.eye(4)
        T[:3, :3] = R
        T[:3, 3] = -R @ (rs.uniform(-max_shift, max_shift, 3) * np.array([1, 0.3, 1]))
        return T
    Ts, Tt = cam(), cam()
    return render_room_scan(Ts, half, tex), render_room_scan(Tt, half, tex), Tt @ np.linalg.inv(Ts)

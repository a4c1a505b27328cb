"""Drop-in for the solver-facing part of the reference's RPModule/rputil.py.

Only ``opts`` (rputil.py:11-22) is on the hot path; it keeps the reference's field
names and defaults because callers mutate the fields freely
(evaluation.py:97-106, trainRelativePoseModuleRecFD.py:254-257)."""
import ctypes

import numpy as np


class opts(object):
    def __init__(self, sigmaAngle1=0.523 / 2, sigmaAngle2=0.523 / 2, sigmaDist=0.08 / 2, sigmaFeat=0.01):
        self.distThre = 0.08
        self.distSepThre = 1.5 * 0.08
        self.angleThre = 45 / 180. * np.pi
        self.sigmaAngle1 = sigmaAngle1
        self.sigmaAngle2 = sigmaAngle2
        self.sigmaDist = sigmaDist
        self.sigmaFeat = sigmaFeat
        self.mu = 0.3
        self.topK = 5
        self.method = 'irls+sm'


def angular_distance_np(R_hat, R):
    """Rotation angle between two (batches of) rotations in degrees (rputil.py:24-35)."""
    R_hat = np.asarray(R_hat).reshape(-1, 3, 3)
    R = np.asarray(R).reshape(-1, 3, 3)
    tr = np.einsum('nij,nij->n', R_hat, R)
    return np.arccos(((tr - 1) / 2).clip(-1, 1)) / np.pi * 180.0


def interpolate(feat, pt):
    """Bilinear descriptor gather (rputil.py:43-58).  feat: torch CUDA float32 [C,H,W]; pt: [K,2] normalised (x,y)
    (torch tensor or numpy).  Returns a torch CUDA tensor [C,K], like the reference.  Runs csrc/scnet.cu:interpolate_kernel."""
    import torch
    from .. import _lib
    lib = _lib.load()
    if not feat.is_cuda:
        raise RuntimeError("relativepose_b200.interpolate needs a CUDA feature map (no CPU fallback)")
    feat = feat.contiguous().float()
    pt = torch.as_tensor(pt, dtype=torch.float32).to(feat.device).contiguous()
    C, H, W = feat.shape
    K = pt.shape[0]
    out = torch.empty((C, K), dtype=torch.float32, device=feat.device)
    with torch.cuda.device(feat.device):
        _lib.check(lib.rp_interpolate(feat.data_ptr(), C, H, W, pt.data_ptr(), K, out.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream), "rp_interpolate")
    return out


_SKYBOX_FACE_R = np.array([
    [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
    [[0, 0, -1], [0, 1, 0], [1, 0, 0]],
    [[-1, 0, 0], [0, 1, 0], [0, 0, -1]],
    [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]], dtype=np.float64)      # rputil.py:66-70 (rotation part)


def getPixel(depth, normal, pts, dataset='suncg', representation='skybox'):
    """Pixel -> 3-D point and normal on the 160x640 four-face skybox (rputil.py:61-119): bilinear depth/normal at the
    sub-pixel location, back-projection through the face's pinhole, face rotation (suncg: face index; scannet and
    matterport: shifted by one).  pts [n,2] pixel (x,y).  Returns (pc [3,n], nn [n,3]) like the reference."""
    assert representation == 'skybox'
    Hf = 160
    assert depth.shape[0] == Hf and depth.shape[1] == 4 * Hf
    pts = np.asarray(pts, dtype=np.float64)
    tp = np.floor(pts).astype('int')
    fx, fy = pts[:, 0] - tp[:, 0], pts[:, 1] - tp[:, 1]
    w00, w01, w10, w11 = (1 - fy) * (1 - fx), fx * (1 - fy), fy * (1 - fx), fx * fy

    def bil(img):
        a, b = img[tp[:, 1], tp[:, 0]], img[tp[:, 1], tp[:, 0] + 1]
        c, d = img[tp[:, 1] + 1, tp[:, 0]], img[tp[:, 1] + 1, tp[:, 0] + 1]
        if img.ndim == 3:
            return a * w00[:, None] + b * w01[:, None] + c * w10[:, None] + d * w11[:, None]
        return a * w00 + b * w01 + c * w10 + d * w11

    val = bil(depth)
    nn = bil(normal)
    nn = nn / np.linalg.norm(nn, axis=1, keepdims=True)
    face = (pts[:, 0] // Hf).astype('int')
    ridx = face if 'suncg' in dataset else (face - 1) % 4
    y = (0.5 - pts[:, 1] / Hf) * 2 * val
    x = ((pts[:, 0] - face * Hf) / Hf - 0.5) * 2 * val
    local = np.stack((x, y, -val), 1)
    pc = np.einsum('nij,nj->ni', _SKYBOX_FACE_R[ridx], local)
    return pc.T, nn


# ---------------------------------------------------------------------------------------------------------------------
# Keypoint detection + augmentation (rputil.py:141-353, Sampling :355-371).
def _sampling_device(fn, n, dev, *args):
    """Launch rp_match_sample / rp_heat_sample on ``dev`` (the device of the caller's tensors, not the current one)."""
    import torch
    from .. import _lib
    with torch.cuda.device(dev):
        out = torch.empty((n, args[-2], 2), dtype=torch.float64, device=dev)
        need = ctypes.c_size_t(0)
        _lib.check(_lib.load().rp_match_sample_workspace_bytes(n, ctypes.byref(need)), "rp_match_sample_workspace_bytes")
        ws = torch.empty((need.value,), dtype=torch.uint8, device=dev)
        _lib.check(fn(*args, out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "rp_*_sample")
    return out.cpu().numpy()


def Sampling(heatmap, K):
    """rputil.Sampling (rputil.py:355-371): heatmap [n,h,w] of squared descriptor distances (numpy or torch) -> [n,K,2]
    (x,y): K rounds of argmax of exp(-d/2) with a 15-pixel suppression window.  Runs csrc/rp_keypoint.cu."""
    import torch
    from .. import _lib
    lib = _lib.load()
    if torch.is_tensor(heatmap) and heatmap.is_cuda:
        d = heatmap.float().contiguous()                     # stays on the caller's device
    else:
        d = torch.as_tensor(np.asarray(heatmap), dtype=torch.float32).cuda().contiguous()
    n, h, w = d.shape
    return _sampling_device(lib.rp_heat_sample, n, d.device, d.data_ptr(), n, h, w, K, 15)


def match_sample(q, feat, K=2):
    """The fused form of ``Sampling(torch_op.npy((q.unsqueeze(2) - feat.view(C,1,-1)).pow(2).sum(0).view(n,H,W)), K)``
    (rputil.py:187-190): q [C,n] CUDA float32 (what ``interpolate`` returns, possibly column-selected), feat [C,H,W]
    CUDA float32 -> [n,K,2] numpy.  The [n,H,W] distance map is never materialised."""
    from .. import _lib
    lib = _lib.load()
    q = q.contiguous().float()
    feat = feat.contiguous().float()
    C, n = q.shape
    q = q.to(feat.device)
    return _sampling_device(lib.rp_match_sample, n, feat.device, q.data_ptr(), C, n, feat.data_ptr(), feat.shape[1], feat.shape[2], K, 15)


def _default_sift(gray):
    import cv2
    mk = cv2.xfeatures2d.SIFT_create if hasattr(cv2, 'xfeatures2d') and hasattr(cv2.xfeatures2d, 'SIFT_create') else cv2.SIFT_create
    kps, _ = mk(contrastThreshold=0.02).detectAndCompute(gray, None)              # rputil.py:152,256
    return np.array([m.pt for m in kps], dtype=np.float64).reshape(-1, 2)


def _keypoints_after_sift(pts, ptt, feats, featt, rng, kinect):
    H, W = 160, 640
    N_SIFT_MATCH, TOPK, MARKER = 30, 2, 0.99
    N_RANDOM = 100 if kinect else 30

    def norm(p):
        qn = p.copy().astype('float')
        qn[:, 0] /= W
        qn[:, 1] /= H
        return qn

    def aug(q, feat):
        a = match_sample(q, feat, TOPK).reshape(-1, 2)
        return a[(a[:, 0] < W - 1) * (a[:, 1] < H - 1)]

    def observed(p):
        if kinect:                                                              # rputil.py:344,348 / :318
            return ((p[:, 0] >= H + H // 2 - 88 // 2) * (p[:, 0] <= H + H // 2 + 88 // 2) *
                    (p[:, 1] >= H // 2 - 66 // 2) * (p[:, 1] <= H // 2 + 66 // 2))
        return (p[:, 0] >= H) * (p[:, 0] <= H * 2)                              # :225,229 / :203
    if kinect:
        pts = pts[rng.choice(range(len(pts)), 300), :]                          # :281-282
        ptt = ptt[rng.choice(range(len(ptt)), 300), :]
    fs0 = interpolate(feats, norm(pts))
    ft0 = interpolate(featt, norm(ptt))
    fsselect = rng.choice(range(pts.shape[0]), min(N_SIFT_MATCH, pts.shape[0]))    # :184-185
    ftselect = rng.choice(range(ptt.shape[0]), min(N_SIFT_MATCH, ptt.shape[0]))
    pttAug = aug(fs0[:, fsselect], featt)                                       # :187-190
    ptsAug = aug(ft0[:, ftselect], feats)
    pts = np.concatenate((pts, ptsAug))
    ptt = np.concatenate((ptt, pttAug))
    n_rand = 120 if kinect else N_RANDOM                                        # :313 / :200
    xs = (rng.rand(n_rand) * W).astype('int').clip(0, W - 2)
    ys = (rng.rand(n_rand) * H).astype('int').clip(0, H - 2)
    ptsrnd = np.stack((xs, ys), 1)
    ptsrnd = ptsrnd[~observed(ptsrnd)]
    fs0 = interpolate(feats, norm(ptsrnd))
    fsselect = rng.choice(range(ptsrnd.shape[0]), min(N_RANDOM, ptsrnd.shape[0]))
    pttAug = aug(fs0[:, fsselect], featt)
    pts = np.concatenate((pts, ptsrnd[fsselect]))
    ptt = np.concatenate((ptt, pttAug))
    ptsW = np.ones(len(pts)); ptsW[~observed(pts)] *= MARKER
    pttW = np.ones(len(ptt)); pttW[~observed(ptt)] *= MARKER
    return pts, norm(pts), ptsW, ptt, norm(ptt), pttW


def getKeypoint(rs, rt, feats, featt, rng=None, sift_fn=None):
    """rputil.getKeypoint (rputil.py:141-236): SIFT keypoints on the observed face of both scans, augmented with the best
    descriptor matches of 30 of them in the other scan's completed feature map and of 30 random unobserved pixels.
    rs/rt uint8 [160,640,3]; feats/featt CUDA float32 [32,160,640].  ``rng``: numpy RandomState (default: the global
    ``numpy.random`` state, like the reference); ``sift_fn(gray) -> [n,2]`` overrides the OpenCV detector.
    Returns (pts, ptsNorm, ptsW, ptt, pttNorm, pttW) or six Nones when a scan has no SIFT keypoint."""
    import cv2
    H = 160
    rng = np.random if rng is None else rng
    sift_fn = sift_fn or _default_sift
    kps = sift_fn(cv2.cvtColor(rs, cv2.COLOR_BGR2GRAY)[:, H:H * 2])
    if not len(kps):
        return None, None, None, None, None, None
    kpt = sift_fn(cv2.cvtColor(rt, cv2.COLOR_BGR2GRAY)[:, H:H * 2])
    if not len(kpt):
        return None, None, None, None, None, None
    pts, ptt = np.asarray(kps, dtype=np.float64).copy(), np.asarray(kpt, dtype=np.float64).copy()
    pts[:, 0] += H
    ptt[:, 0] += H
    return _keypoints_after_sift(pts, ptt, feats, featt, rng, False)


def getKeypoint_kinect(rs, rt, feats, featt, rs_full, rt_full, rng=None, sift_fn=None):
    """rputil.getKeypoint_kinect (rputil.py:239-353): SIFT on the original 480x640 frames mapped into the 66x88 window of
    the panorama, 300 of them sub-sampled, then the same augmentation with 100 random points."""
    import cv2
    H = 160
    rng = np.random if rng is None else rng
    sift_fn = sift_fn or _default_sift

    def place(kp):
        p = np.asarray(kp, dtype=np.float64).copy()
        p[:, 0] = p[:, 0] / 640 * 88
        p[:, 1] = p[:, 1] / 480 * 66
        p[:, 0] += H + H // 2 - 88 // 2
        p[:, 1] += H // 2 - 66 // 2
        return p
    kps = sift_fn(cv2.cvtColor(rs_full, cv2.COLOR_BGR2GRAY))
    if not len(kps):
        return None, None, None, None, None, None
    kpt = sift_fn(cv2.cvtColor(rt_full, cv2.COLOR_BGR2GRAY))
    if not len(kpt):
        return None, None, None, None, None, None
    return _keypoints_after_sift(place(kps), place(kpt), feats, featt, rng, True)

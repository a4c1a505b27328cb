"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's keypoint augmentation.

Follows /root/reference/RPModule/rputil.py: ``getKeypoint`` (:141-236), ``getKeypoint_kinect`` (:239-353) downstream of
the SIFT detector, ``Sampling`` (:355-371) and ``interpolate`` (:43-58).  The SIFT detections themselves (OpenCV, CPU) are
an input here: the golden generator records what ``cv2.SIFT_create(contrastThreshold=0.02).detectAndCompute`` returned.
Random draws come from the ``rng`` argument (a ``numpy.random.RandomState``; the reference uses the global state) in the
reference's call order: choice(source), choice(target), rand(xs), rand(ys), choice(random points) -- and, for the Kinect
variant, the two N_SIFT sub-samplings first.

Only tests/ may import this module.  Pinned by tests/golden/keypoint_golden.npz (outputs of the unmodified reference).
"""
import numpy as np

H, W = 160, 640


def interpolate(feat, pt):
    """rputil.py:43-58 in float32.  feat [C,H,W] float32, pt [K,2] normalised -> [C,K] float32."""
    feat = np.asarray(feat, dtype=np.float32)
    pt = np.asarray(pt, dtype=np.float32)
    h, w = feat.shape[1], feat.shape[2]
    x = pt[:, 0] * np.float32(w - 1)
    y = pt[:, 1] * np.float32(h - 1)
    x0 = np.floor(x)
    y0 = np.floor(y)
    xi, yi = x0.astype(np.int64), y0.astype(np.int64)
    one = np.float32(1)
    return (feat[:, yi, xi] * (x0 + one - x) * (y0 + one - y) + feat[:, yi + 1, xi] * (x0 + one - x) * (y - y0) +
            feat[:, yi, xi + 1] * (x - x0) * (y0 + one - y) + feat[:, yi + 1, xi + 1] * (x - x0) * (y - y0))


def dense_dist(q, feat):
    """``(q.unsqueeze(2) - feat.view(C,1,-1)).pow(2).sum(0)`` (rputil.py:187,189,209): q [C,n], feat [C,H,W] -> [n,H,W]
    float32, channels accumulated in order."""
    C = feat.shape[0]
    f = np.asarray(feat, dtype=np.float32).reshape(C, 1, -1)
    q = np.asarray(q, dtype=np.float32)
    acc = np.zeros((q.shape[1], f.shape[2]), np.float32)
    for c in range(C):
        d = q[c][:, None] - f[c]
        acc = acc + d * d
    return acc.reshape(q.shape[1], feat.shape[1], feat.shape[2])


def sampling(heatmap, K):
    """rputil.py:355-371: K rounds of argmax + suppression of a 15-pixel window (set to the map's current minimum)."""
    heatmap = np.exp(-heatmap / 2)
    n, h, w = heatmap.shape
    pt = np.zeros([n, K, 2])
    WINDOW_SZ = 15
    for i in range(n):
        for j in range(K):
            idx = np.argmax(heatmap[i])
            coord = np.unravel_index(idx, heatmap[i].shape)[::-1]
            pt[i, j, :] = coord
            topl = [max(0, coord[0] - WINDOW_SZ), max(0, coord[1] - WINDOW_SZ)]
            botr = [min(w - 1, coord[0] + WINDOW_SZ), min(h - 1, coord[1] + WINDOW_SZ)]
            heatmap[i][topl[1]:botr[1], topl[0]:botr[0]] = heatmap[i].min()
    return pt


def _norm(p):
    q = p.copy().astype('float')
    q[:, 0] /= W
    q[:, 1] /= H
    return q


def _aug(q, feat, K):
    a = sampling(dense_dist(q, feat), K).reshape(-1, 2)
    return a[(a[:, 0] < W - 1) * (a[:, 1] < H - 1)]


def get_keypoint(pts, ptt, feats, featt, rng, kinect=False):
    """Everything of getKeypoint (:166-236) / getKeypoint_kinect (:285-353) after the SIFT points were placed in panorama
    coordinates.  pts/ptt [n,2] float64."""
    N_SIFT_MATCH, TOPK, MARKER = 30, 2, 0.99
    N_RANDOM = 100 if kinect else 30
    if kinect:
        pts = pts[rng.choice(range(len(pts)), 300), :]                      # :281-282
        ptt = ptt[rng.choice(range(len(ptt)), 300), :]
    fs0 = interpolate(feats, _norm(pts))
    ft0 = interpolate(featt, _norm(ptt))
    fsselect = rng.choice(range(pts.shape[0]), min(N_SIFT_MATCH, pts.shape[0]))
    ftselect = rng.choice(range(ptt.shape[0]), min(N_SIFT_MATCH, ptt.shape[0]))
    pttAug = _aug(fs0[:, fsselect], featt, TOPK)
    ptsAug = _aug(ft0[:, ftselect], feats, TOPK)
    pts = np.concatenate((pts, ptsAug))
    ptt = np.concatenate((ptt, pttAug))
    n_rand = 120 if kinect else N_RANDOM                                      # :313 (N=120) / :200
    xs = (rng.rand(n_rand) * W).astype('int').clip(0, W - 2)
    ys = (rng.rand(n_rand) * H).astype('int').clip(0, H - 2)
    ptsrnd = np.stack((xs, ys), 1)

    def observed(p):
        if kinect:
            return ((p[:, 0] >= H + H // 2 - 88 // 2) * (p[:, 0] <= H + H // 2 + 88 // 2) *
                    (p[:, 1] >= H // 2 - 66 // 2) * (p[:, 1] <= H // 2 + 66 // 2))
        return (p[:, 0] >= H) * (p[:, 0] <= H * 2)
    ptsrnd = ptsrnd[~observed(ptsrnd)]
    fs0 = interpolate(feats, _norm(ptsrnd))
    fsselect = rng.choice(range(ptsrnd.shape[0]), min(N_RANDOM, ptsrnd.shape[0]))
    pttAug = _aug(fs0[:, fsselect], featt, TOPK)
    pts = np.concatenate((pts, ptsrnd[fsselect]))
    ptt = np.concatenate((ptt, pttAug))
    ptsW = np.ones(len(pts)); ptsW[~observed(pts)] *= MARKER
    pttW = np.ones(len(ptt)); pttW[~observed(ptt)] *= MARKER
    return pts, _norm(pts), ptsW, ptt, _norm(ptt), pttW


def place_sift(kp, kinect=False):
    """SIFT detections -> panorama pixel coordinates (:160-161 / :262-265)."""
    p = np.asarray(kp, dtype=np.float64).copy()
    if kinect:
        p[:, 0] = p[:, 0] / 640 * 88
        p[:, 1] = p[:, 1] / 480 * 66
        p[:, 0] += H + H // 2 - 88 // 2
        p[:, 1] += H // 2 - 66 // 2
    else:
        p[:, 0] += H
    return p

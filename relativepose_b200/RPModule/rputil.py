"""Drop-in for the solver-facing part of the reference's RPModule/rputil.py.

Only ``opts`` (rputil.py:11-22) is on the hot path; it keeps the reference's field
names and defaults because callers mutate the fields freely
(evaluation.py:97-106, trainRelativePoseModuleRecFD.py:254-257)."""
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

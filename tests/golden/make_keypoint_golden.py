"""Golden vectors for rputil.getKeypoint / getKeypoint_kinect (RPModule/rputil.py:141-353) from the UNMODIFIED reference on
CPU: cv2.xfeatures2d.SIFT_create is aliased to cv2.SIFT_create (same detector, the contrib namespace is gone in OpenCV
4.13) and torch_op.v keeps tensors on the CPU.  Stored: the SIFT detections (an input of the restatement), the numpy seed,
the six outputs.  Images / feature maps are regenerated from seeds (relativepose_b200/synth.py)."""
import os, sys, types
import numpy as np, torch, cv2
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_loader import load_reference_rpmodule  # noqa: E402
from relativepose_b200 import synth  # noqa: E402

ru = load_reference_rpmodule()._rputil
cv2.xfeatures2d = types.SimpleNamespace(SIFT_create=cv2.SIFT_create)
ru.torch_op.v = lambda var, cuda=True, volatile=False: (torch.from_numpy(var).float() if isinstance(var, np.ndarray) else var.float())
sift = cv2.SIFT_create(contrastThreshold=0.02)


def detect(gray):
    kps, _ = sift.detectAndCompute(gray, None)
    return np.array([m.pt for m in kps], dtype=np.float64).reshape(-1, 2)


blob, names = {}, []
for name, kinect, seed in (("suncg_a", 0, 1), ("suncg_b", 0, 5), ("kinect_a", 1, 9)):
    rs_, rt_ = synth.make_texture_image(seed), synth.make_texture_image(seed + 100)
    fs, ft = synth.make_feature_map(seed + 200), synth.make_feature_map(seed + 300)
    np.random.seed(seed + 400)
    if kinect:
        rsf, rtf = synth.make_texture_image(seed + 500, 480, 640), synth.make_texture_image(seed + 600, 480, 640)
        out = ru.getKeypoint_kinect(rs_, rt_, torch.from_numpy(fs), torch.from_numpy(ft), rsf, rtf)
        kps, kpt = detect(cv2.cvtColor(rsf, cv2.COLOR_BGR2GRAY)), detect(cv2.cvtColor(rtf, cv2.COLOR_BGR2GRAY))
    else:
        out = ru.getKeypoint(rs_, rt_, torch.from_numpy(fs), torch.from_numpy(ft))
        kps = detect(cv2.cvtColor(rs_, cv2.COLOR_BGR2GRAY)[:, 160:320])
        kpt = detect(cv2.cvtColor(rt_, cv2.COLOR_BGR2GRAY)[:, 160:320])
    names.append(name)
    blob[name + '/meta'] = np.array([kinect, seed])
    blob[name + '/kps'], blob[name + '/kpt'] = kps, kpt
    for k, o in zip(('pts', 'ptsNorm', 'ptsW', 'ptt', 'pttNorm', 'pttW'), out):
        blob[name + '/' + k] = o
    print(name, kps.shape, kpt.shape, [o.shape for o in out])
blob['names'] = np.array(names)
np.savez_compressed(os.path.join(HERE, 'keypoint_golden.npz'), **blob)
print("ok", os.path.getsize(os.path.join(HERE, 'keypoint_golden.npz')))

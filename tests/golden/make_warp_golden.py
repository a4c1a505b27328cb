"""Golden vectors for util.warping (util.py:94-172) from the UNMODIFIED reference, executed on CPU through
oracle/ref_loader.py (the reference solver module imports util).  Inputs are regenerated from seeds
(relativepose_b200/synth.py: make_warp_view / make_pose); stored per case: the validity mask (packed bits), the
per-channel float64 sums and the full values at every 7th written pixel."""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_loader import load_reference_rpmodule  # noqa: E402
from relativepose_b200 import synth  # noqa: E402

util = load_reference_rpmodule().util
blob, names = {}, []
for ds in ('suncg', 'matterport', 'scannet'):
    for seed in (0, 1):
        for inv in (0, 1):
            view = synth.make_warp_view(seed, ds)
            R = synth.make_pose(seed)
            if inv:
                R = np.linalg.inv(R)
            out = util.warping(view, R, ds)                       # [1,8,160,640] float64
            name = "%s_s%d_%s" % (ds, seed, "inv" if inv else "fwd")
            names.append(name)
            m = out[0, 7] != 0
            idx = np.flatnonzero(m.ravel())[::7]
            blob[name + '/meta'] = np.array([seed, inv])
            blob[name + '/dataset'] = np.array(ds)
            blob[name + '/mask'] = np.packbits(m.ravel())
            blob[name + '/sums'] = out[0].reshape(8, -1).sum(1)
            blob[name + '/idx'] = idx.astype(np.int32)
            blob[name + '/vals'] = out[0].reshape(8, -1)[:, idx].astype(np.float32)
            print(name, int(m.sum()), len(idx))
# Pano2PointCloud (util.py:751-811) on one depth map per dataset: checksums + a strided sample
d = synth.make_warp_view(5, 'suncg')[0, 6].copy()
full = np.random.RandomState(9).uniform(0.5, 5, (160, 640)).astype(np.float32)
full[np.random.RandomState(10).rand(160, 640) < 0.05] = 0
for ds in ('suncg', 'matterport', 'scannet'):
    pc = util.Pano2PointCloud(full, ds)
    blob['pano_' + ds + '/shape'] = np.array(pc.shape)
    blob['pano_' + ds + '/sums'] = pc.sum(1)
    blob['pano_' + ds + '/sample'] = pc[:, ::97]
blob['names'] = np.array(names)
np.savez_compressed(os.path.join(HERE, 'warp_golden.npz'), **blob)
print("ok", os.path.getsize(os.path.join(HERE, 'warp_golden.npz')))

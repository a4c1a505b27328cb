"""Golden vectors for rputil.interpolate / rputil.getPixel from the reference's own functions (RPModule/rputil.py:43-119),
executed on CPU through oracle/ref_loader.py."""
import os, sys
import numpy as np
import torch
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.ref_loader import load_reference_rpmodule  # noqa: E402

ref = load_reference_rpmodule()._rputil
rs = np.random.RandomState(5)
feat = rs.randn(32, 160, 640).astype(np.float32)
ptn = rs.uniform(0.01, 0.98, size=(77, 2)).astype(np.float32)
with torch.no_grad():
    val = ref.interpolate(torch.from_numpy(feat), torch.from_numpy(ptn)).numpy()
depth = rs.uniform(0.5, 5, size=(160, 640))
normal = rs.randn(160, 640, 3); normal /= np.linalg.norm(normal, axis=2, keepdims=True)
pts = np.stack((rs.uniform(0, 638, 60), rs.uniform(0, 158, 60)), 1)
blob = {'feat_seed': np.array(5), 'ptn': ptn, 'interp': val, 'depth': depth, 'normal': normal, 'pts': pts}
for ds in ('suncg', 'matterport'):
    pc, nn = ref.getPixel(depth, normal, pts, dataset=ds)
    blob['pc_' + ds], blob['nn_' + ds] = pc, nn
np.savez_compressed(os.path.join(HERE, 'rputil_golden.npz'), **blob)
print("ok", val.shape, blob['pc_suncg'].shape)

"""Golden vectors for the SCNet constructor variants (mymodel.py:15-39 ``batchnorm=0``: biased convolutions without
BatchNorm; :333-357 ``skipLayer=0``: decoder without the encoder concatenations; a partial ``outputType``), from the UNMODIFIED
reference ``SCNet`` class on CPU -- same procedure as make_scnet_golden.py (the reference module gets the ``state_dict`` of this
repo's container built under ``torch.manual_seed(0)``; the oracle is checked against it here).  Note: the reference's own
forward fails for skipLayer=0 with any of the rgb / n / d heads (deconv1rgb/n/d are built for 64 input channels, :190,198,206,
but get 32, :341-353), so skipLayer=0 is only exercised with heads drawn from 's', 'f'.  Stored: output sub-sampled every
8th row / 16th column + per-channel moments."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import scnet_oracle  # noqa: E402
from relativepose_b200 import synth  # noqa: E402
from relativepose_b200.model.mymodel import SCNet  # noqa: E402

VARIANTS = (
    # name, batchnorm, skipLayer, outputType, snumclass, useTanh, dataset, input seed
    ("nobn", 0, 1, 'rgbdnsf', 15, 1, 'suncg', 2),
    ("noskip_sf", 1, 0, 'sf', 21, 0, 'scannet', 3),
    ("partial_df", 1, 1, 'df', 15, 1, 'suncg', 4),
    ("nobn_noskip_f", 0, 0, 'f', 15, 1, 'suncg', 5),
)


def main():
    torch.set_num_threads(8)
    spec = importlib.util.spec_from_file_location("ref_mymodel", "/root/reference/model/mymodel.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    blob = {}
    for name, bn, skip, otype, snum, tanh, ds, seed in VARIANTS:
        a = types.SimpleNamespace(batchnorm=bn, useTanh=tanh, skipLayer=skip, outputType=otype, snumclass=snum)
        torch.manual_seed(0)
        mine = SCNet(a)
        sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
        net = ref.SCNet(a)
        assert list(net.state_dict().keys()) == list(sd.keys()), "state_dict keys differ from the reference"
        net.load_state_dict(sd)
        x = torch.from_numpy(synth.make_panorama_pair(seed, ds))
        with torch.no_grad():
            y_ref = net(x)
            y_or = scnet_oracle.forward(sd, x, snum, bool(tanh), skip=bool(skip), heads=tuple(mine.heads))
        err = (y_ref - y_or).abs().max().item()
        print("%s: out %s  |oracle-ref|max = %.3e   |y|max = %.3f" % (name, tuple(y_ref.shape), err, y_ref.abs().max().item()))
        assert err <= 1e-4 * max(1.0, y_ref.abs().max().item())
        y = y_ref.numpy()
        blob[name + '/sub'] = y[:, :, ::8, ::16].copy()
        blob[name + '/mean'] = y.mean(axis=(2, 3))
        blob[name + '/std'] = y.std(axis=(2, 3))
        blob[name + '/meta'] = np.array([bn, skip, snum, tanh, seed, float(sum(v.double().abs().sum().item() for v in sd.values()))])
        blob[name + '/otype'] = np.array(otype)
        blob[name + '/dataset'] = np.array(ds)
    blob['names'] = np.array([v[0] for v in VARIANTS])
    path = os.path.join(HERE, "scnet_variants_golden.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

"""Generate tests/golden/scnet_golden.npz by running the UNMODIFIED reference ``SCNet`` (model/mymodel.py) on CPU.

The reference module gets the ``state_dict`` of ``relativepose_b200.model.mymodel.SCNet`` built with
``torch.manual_seed(0)`` (same key names and shapes -- that is itself asserted), so the weights are identical by
construction; the oracle (oracle/scnet_oracle.py) is checked against the reference here, and the output is stored
sub-sampled (every 4th row, 8th column) together with per-channel moments of the full output.
"""
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


def load_reference_model_module():
    spec = importlib.util.spec_from_file_location("ref_mymodel", "/root/reference/model/mymodel.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def args_for(snumclass, use_tanh=1):
    a = types.SimpleNamespace()
    a.batchnorm, a.useTanh, a.skipLayer, a.outputType, a.snumclass = 1, use_tanh, 1, 'rgbdnsf', snumclass
    return a


def state_checksum(sd):
    return float(sum(v.double().abs().sum().item() for v in sd.values()))


def main():
    torch.set_num_threads(8)
    ref = load_reference_model_module()
    blob = {}
    for name, ds, snum, tanh, seed in (("suncg", "suncg", 15, 1, 0), ("scannet", "scannet", 21, 0, 1)):
        torch.manual_seed(0)
        mine = SCNet(args_for(snum, tanh))
        sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
        net = ref.SCNet(args_for(snum, tanh))
        assert list(net.state_dict().keys()) == list(sd.keys()), "state_dict keys differ from the reference"
        for k, v in net.state_dict().items():
            assert v.shape == sd[k].shape, k
        net.load_state_dict(sd)
        x = torch.from_numpy(synth.make_panorama_pair(seed, ds))
        with torch.no_grad():
            y_ref = net(x)
            y_or = scnet_oracle.forward(sd, x, snum, bool(tanh))
        err = (y_ref - y_or).abs().max().item()
        print("%s: out %s  |oracle-ref|max = %.3e   |y|max = %.3f" % (name, tuple(y_ref.shape), err, y_ref.abs().max().item()))
        assert err <= 1e-5
        y = y_ref.numpy()
        blob[name + '/sub'] = y[:, :, ::4, ::8].copy()
        blob[name + '/mean'] = y.mean(axis=(2, 3))
        blob[name + '/std'] = y.std(axis=(2, 3))
        blob[name + '/meta'] = np.array([snum, tanh, seed, state_checksum(sd)], dtype=np.float64)
        blob[name + '/dataset'] = np.array(ds)
    path = os.path.join(HERE, "scnet_golden.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

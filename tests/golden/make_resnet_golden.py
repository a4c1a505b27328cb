"""Generate tests/golden/resnet_golden.npz from the reference ``Resnet18_8s`` class (model/mymodel.py:41-122).

The class needs a forked torchvision (README.md:11) that is not vendored; as SURVEY.md section 8c describes, a shim that
drops the fork-only kwargs over stock ``torchvision.models.resnet18(weights=None)`` constructs and runs it (the forward
only touches the stock trunk).  Fork-specific behaviour therefore stays UNPINNED; what is pinned is mymodel.py's own
forward on the stock trunk.  Weights: the state_dict of relativepose_b200.model.mymodel.Resnet18_8s (seed 0)."""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import resnet_oracle  # noqa: E402
from relativepose_b200.model.mymodel import Resnet18_8s  # noqa: E402


def main():
    torch.set_num_threads(8)
    stock = torchvision.models.resnet18
    torchvision.models.resnet18 = lambda **kw: stock(weights=None)       # drops fully_conv/pretrained/output_stride/...
    spec = importlib.util.spec_from_file_location("ref_mymodel", "/root/reference/model/mymodel.py")
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    blob = {}
    for name, tanh, n in (("pair_tanh", 1, 2), ("batch4_notanh", 0, 4)):
        args = types.SimpleNamespace(num_input=7, useTanh=tanh)
        torch.manual_seed(0)
        mine = Resnet18_8s(args)
        sd = {k: v.detach().clone() for k, v in mine.state_dict().items()}
        net = ref.Resnet18_8s(args)
        assert list(net.state_dict().keys()) == list(sd.keys()), "state_dict keys differ from the reference"
        net.load_state_dict(sd)
        net.train()
        rs = np.random.RandomState(7 + n)
        x = torch.from_numpy(rs.uniform(-1, 1, size=(n, 7, 160, 640)).astype(np.float32))
        with torch.no_grad():
            y_ref = net(x)
        y_or = resnet_oracle.forward(sd, x, bool(tanh))
        err = (y_ref - y_or).abs().max().item()
        print("%s: out %s |oracle-ref|max = %.3e |y|max %.3f" % (name, tuple(y_ref.shape), err, y_ref.abs().max().item()))
        assert err <= 1e-5
        y = y_ref.numpy()
        blob[name + '/sub'] = y[:, :, ::4, ::8].copy()
        blob[name + '/meta'] = np.array([tanh, n, 7 + n, float(sum(v.double().abs().sum().item() for v in sd.values()))])
    path = os.path.join(HERE, "resnet_golden.npz")
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

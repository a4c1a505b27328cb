"""TEST INFRASTRUCTURE ONLY -- plain PyTorch fp32 restatement of the reference's ``Resnet18_8s.forward``
(model/mymodel.py:82-122) driven by a ``state_dict`` with the reference's key names.

The trunk is the stock torchvision ResNet-18 (conv1, bn1, relu, maxpool, layer1..4 of BasicBlocks); the reference
builds it from a forked torchvision that is not vendored (README.md:11), so fork-specific behaviour is UNPINNED
(SURVEY.md section 8c).  tests/golden/make_resnet_golden.py pins this file against the reference class itself with a
kwarg-dropping shim over stock ``torchvision.models.resnet18``.  BatchNorm uses batch statistics of the call
(the module is never put in eval mode: mainPanoCompletion2view.py:132,268-274)."""
import torch
import torch.nn.functional as F

EPS = 1e-5


def _bn(sd, name, y):
    return F.batch_norm(y, None, None, sd[name + '.weight'], sd[name + '.bias'], True, 0.0, EPS)


def _block(sd, pre, x, stride, trace=None):
    y = F.relu(_bn(sd, pre + '.bn1', F.conv2d(x, sd[pre + '.conv1.weight'], None, stride, 1)))
    y = _bn(sd, pre + '.bn2', F.conv2d(y, sd[pre + '.conv2.weight'], None, 1, 1))
    idn = x
    if pre + '.downsample.0.weight' in sd:
        idn = _bn(sd, pre + '.downsample.1', F.conv2d(x, sd[pre + '.downsample.0.weight'], None, stride, 0))
    out = F.relu(y + idn)
    if trace is not None:
        trace[pre] = out
    return out


def forward(sd, x, use_tanh=True, trace=None):
    with torch.no_grad():
        size = x.shape[2:]
        t = 'resnet18_32s.'
        y = F.relu(_bn(sd, t + 'bn1', F.conv2d(x, sd[t + 'conv1.weight'], None, 2, 3)))        # mymodel.py:85-87
        y = F.max_pool2d(y, 3, 2, 1)                                                            # :88
        if trace is not None:
            trace['pool'] = y
        for li, stride in ((1, 1), (2, 2), (3, 2), (4, 2)):
            y = _block(sd, t + 'layer%d.0' % li, y, stride, trace)
            y = _block(sd, t + 'layer%d.1' % li, y, 1, trace)
            if li == 2:
                l8 = F.conv2d(y, sd['score_8s.weight'], sd['score_8s.bias'])                    # :94
            if li == 3:
                l16 = F.conv2d(y, sd['score_16s.weight'], sd['score_16s.bias'])                 # :97
        l32 = F.conv2d(y, sd['score_32s.weight'], sd['score_32s.bias'])                         # :100
        l16 = l16 + F.interpolate(l32, size=l16.shape[2:], mode='bilinear', align_corners=False)   # :105-106
        l8 = l8 + F.interpolate(l16, size=l8.shape[2:], mode='bilinear', align_corners=False)      # :108-109
        out = F.interpolate(l8, size=list(size), mode='bilinear', align_corners=False)             # :111
        if use_tanh:
            out = torch.tanh(out)                                                                  # :120-121
        return out

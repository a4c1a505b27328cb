"""Drop-in for the reference's ``model/mymodel.py`` model entry points, backed by the CUDA library.

``SCNet(args)`` keeps the constructor arguments (``batchnorm, useTanh, skipLayer, outputType, snumclass``,
mymodel.py:142-148) and -- because ``net.load_state_dict(checkpoint['state_dict'])`` is part of the interface
(evaluation.py:152-153) -- the exact parameter names and shapes of the reference (103 keys: ``conv1rgb.0.weight``,
``conv1rgb.1.weight``, ``conv1rgb.1.bias``, ..., ``deconv1f.weight/bias``).  The torch sub-modules are parameter
containers only: ``forward`` never calls them; it hands the parameter tensors to the sm_100a kernels in
``csrc/scnet.cu`` (no cuDNN, no torch ops on the hot path).

``weights_init`` restates mymodel.py:6-13 (Xavier-normal conv weights, BN gamma ~ N(1, 0.02), beta = 0) so that
seeded synthetic weights are well conditioned when no checkpoint is available.
"""
import torch
import torch.nn as nn


def weights_init(m):
    """mymodel.py:6-13."""
    name = m.__class__.__name__
    if name.find('Conv') != -1:
        nn.init.xavier_normal_(m.weight.data)
    elif name.find('BatchNorm') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def _block(transposed, cin, cout, k, s, p):
    conv = (nn.ConvTranspose2d if transposed else nn.Conv2d)(cin, cout, kernel_size=k, stride=s, padding=p, bias=False)
    return nn.Sequential(conv, nn.BatchNorm2d(cout, track_running_stats=False), nn.LeakyReLU(0.1, inplace=True))


# (name, transposed, cin, cout, k, s, p) in the reference's registration order (mymodel.py:151-231)
def _layer_table(snumclass, skip):
    ngf, m = 64, (2 if skip else 1)
    t = [('conv1rgb', 0, 4, 32, 3, 1, 1), ('conv2rgb', 0, 32, 64, 4, 2, 1), ('conv3rgb', 0, 64, 128, 4, 2, 1),
         ('conv1n', 0, 4, 32, 3, 1, 1), ('conv2n', 0, 32, 64, 4, 2, 1), ('conv3n', 0, 64, 128, 4, 2, 1),
         ('conv1d', 0, 2, 32, 3, 1, 1), ('conv2d', 0, 32, 64, 4, 2, 1), ('conv3d', 0, 64, 128, 4, 2, 1),
         ('conv4', 0, ngf * 2 * 6, ngf * 4, 4, 2, 1), ('conv5', 0, ngf * 4, ngf * 8, 4, 2, 1),
         ('conv6', 0, ngf * 8, ngf * 8, 4, 2, 1), ('conv7', 0, ngf * 8, ngf * 8, 3, 2, 0),
         ('conv8', 0, ngf * 8, ngf * 8, 3, 1, 1), ('conv9', 0, ngf * 8, ngf * 16, 3, 1, 0),
         ('deconv9', 1, ngf * 16, ngf * 8, 3, 1, 0), ('deconv8', 1, ngf * 8 * m, ngf * 8, 3, 1, 1),
         ('deconv7', 1, ngf * 8 * m, ngf * 8, 3, 2, 0), ('deconv6', 1, ngf * 8 * m, ngf * 8, 4, 2, 1),
         ('deconv5', 1, ngf * 8 * m, ngf * 4, 4, 2, 1), ('deconv4', 1, ngf * 4 * m, ngf * 2, 4, 2, 1)]
    return t


class SCNet(nn.Module):
    def __init__(self, args):
        super(SCNet, self).__init__()
        if not args.batchnorm:
            raise NotImplementedError("relativepose_b200.SCNet implements the batchnorm=1 configuration the reference ships")
        if 'k' in args.outputType:
            raise NotImplementedError("outputType 'k' is dead code in the reference (mymodel.py:327-331 reads undefined tensors)")
        for need in ('rgb', 'n', 'd', 's', 'f'):
            if need not in args.outputType:
                raise NotImplementedError("relativepose_b200.SCNet implements outputType 'rgbdnsf' (evaluation.py:52)")
        if not args.skipLayer:
            raise NotImplementedError("relativepose_b200.SCNet implements skipLayer=1 (opts.py:28 default)")
        self.useTanh = args.useTanh
        self.skipLayer = args.skipLayer
        self.outputType = args.outputType
        self.snumclass = args.snumclass
        ngf, m = 64, 2
        for name, tr, cin, cout, k, s, p in _layer_table(args.snumclass, True):
            setattr(self, name, _block(bool(tr), cin, cout, k, s, p))
        for st, nout in (('rgb', 3), ('n', 3), ('d', 1)):
            setattr(self, 'deconv3' + st, _block(True, ngf * 2 * m, ngf, 4, 2, 1))
            setattr(self, 'deconv2' + st, _block(True, ngf * m, ngf // 2, 4, 2, 1))
            setattr(self, 'deconv1' + st, nn.Conv2d(ngf, nout, 1, 1, 0))
        for st, nout in (('s', args.snumclass), ('f', 32)):
            setattr(self, 'deconv3' + st, _block(True, ngf * 2, ngf, 4, 2, 1))
            setattr(self, 'deconv2' + st, _block(True, ngf, ngf, 4, 2, 1))
            setattr(self, 'deconv1' + st, nn.Conv2d(ngf, nout, 1, 1, 0))
        self.apply(weights_init)
        self._engine = None

    def forward(self, x):
        """x: [2P,16,H,W] float32 CUDA (NCHW, as evaluation.py:242 builds it) -> [2P, 7+snumclass+32, H, W].
        Consecutive image pairs are independent BN groups, exactly like P separate calls of the reference."""
        from .. import scnet_engine
        if self._engine is None:
            self._engine = scnet_engine.ScnetEngine(self)
        return self._engine.forward(x)


# --------------------------------------------------------------------------------------------------------------
# Resnet18_8s (mymodel.py:41-122).  The reference builds its trunk from a *forked* torchvision
# (README.md:11: warmspringwinds/vision, kwargs fully_conv/output_stride/remove_avg_pool_layer) that is not
# vendored; its forward touches only conv1, bn1, relu, maxpool, layer1..4 of the stock ResNet-18 (mymodel.py:85-99).
# The containers below reproduce the stock parameter/buffer names so checkpoints of the reference load
# (``resnet18_32s.layer2.0.downsample.0.weight`` ...).  forward() runs the CUDA kernels; the torch sub-modules are
# never called.
class _BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, cout, stride):
        super(_BasicBlock, self).__init__()
        self.conv1 = nn.Conv2d(cin, cout, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(cout, cout, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.stride = stride
        self.downsample = None
        if stride != 1 or cin != cout:
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride, bias=False), nn.BatchNorm2d(cout))


class _ResNet18Trunk(nn.Module):
    def __init__(self, num_input):
        super(_ResNet18Trunk, self).__init__()
        self.conv1 = nn.Conv2d(num_input, 64, kernel_size=7, stride=2, padding=3, bias=False)     # mymodel.py:57
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = nn.Sequential(_BasicBlock(64, 64, 1), _BasicBlock(64, 64, 1))
        self.layer2 = nn.Sequential(_BasicBlock(64, 128, 2), _BasicBlock(128, 128, 1))
        self.layer3 = nn.Sequential(_BasicBlock(128, 256, 2), _BasicBlock(256, 256, 1))
        self.layer4 = nn.Sequential(_BasicBlock(256, 512, 2), _BasicBlock(512, 512, 1))
        self.fc = nn.Sequential()                                                               # mymodel.py:61


class Resnet18_8s(nn.Module):
    def __init__(self, args):
        super(Resnet18_8s, self).__init__()
        self.args = args
        self.resnet18_32s = _ResNet18Trunk(args.num_input)
        self.score_32s = nn.Conv2d(512, 32, kernel_size=1)
        self.score_16s = nn.Conv2d(256, 32, kernel_size=1)
        self.score_8s = nn.Conv2d(128, 32, kernel_size=1)
        self._engine = None

    def forward(self, x):
        """x: [n,num_input,H,W] float32 CUDA -> [n,32,H,W].  BatchNorm uses the statistics of this call's n images
        (the reference never puts the module in eval mode: mainPanoCompletion2view.py:132,268-274); running-stat
        buffers are not updated."""
        from .. import resnet_engine
        if self._engine is None:
            self._engine = resnet_engine.ResnetEngine(self)
        return self._engine.forward(x)


class segmentation_layer(nn.Module):
    """mymodel.py:126-139: 1x1 head on the 32-d feature map (training-time auxiliary; plain torch)."""

    def __init__(self, args):
        super(segmentation_layer, self).__init__()
        self.segm_layer = nn.Conv2d(32, args.snumclass, kernel_size=1)

    def forward(self, featMap):
        return self.segm_layer(featMap)

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


def _block(transposed, cin, cout, k, s, p, batchnorm=True):
    """conv2d / deconv2d of mymodel.py:15-39: conv (no bias) + BatchNorm (batch statistics) + LeakyReLU(0.1), or -- batchnorm=0 --
    conv with bias + LeakyReLU(0.1).  Same Sequential indices as the reference, hence the same state_dict keys."""
    conv = (nn.ConvTranspose2d if transposed else nn.Conv2d)(cin, cout, kernel_size=k, stride=s, padding=p, bias=not batchnorm)
    if batchnorm:
        return nn.Sequential(conv, nn.BatchNorm2d(cout, track_running_stats=False), nn.LeakyReLU(0.1, inplace=True))
    return nn.Sequential(conv, nn.LeakyReLU(0.1, inplace=True))


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


HEAD_CHANNELS = {'rgb': 3, 'n': 3, 'd': 1, 's': None, 'f': 32}      # 's': args.snumclass
HEAD_ORDER = ('rgb', 'n', 'd', 's', 'f')                             # order of the output channels (mymodel.py:309-377)


class SCNet(nn.Module):
    def __init__(self, args):
        super(SCNet, self).__init__()
        if 'k' in args.outputType:
            raise NotImplementedError("outputType 'k' is dead code in the reference (mymodel.py:327-331 reads undefined tensors)")
        self.heads = [h for h in HEAD_ORDER if (h in args.outputType if h != 'rgb' else 'rgb' in args.outputType)]
        if not self.heads:
            raise ValueError("outputType %r selects no output head" % (args.outputType,))
        if not args.skipLayer and any(h in ('rgb', 'n', 'd') for h in self.heads):
            # the reference's own forward fails there: deconv1rgb/n/d are built for 64 input channels (mymodel.py:190,198,206)
            # but the skipLayer=0 branch feeds them the 32 channels of deconv2* alone (:341,347,353)
            raise NotImplementedError("skipLayer=0 works in the reference only with outputType drawn from 's', 'f'")
        self.batchnorm = int(bool(args.batchnorm))
        self.useTanh = args.useTanh
        self.skipLayer = args.skipLayer
        self.outputType = args.outputType
        self.snumclass = args.snumclass
        ngf, m, bnf = 64, (2 if args.skipLayer else 1), bool(args.batchnorm)
        for name, tr, cin, cout, k, s, p in _layer_table(args.snumclass, bool(args.skipLayer)):
            setattr(self, name, _block(bool(tr), cin, cout, k, s, p, bnf))
        for st in self.heads:                                   # registration order of the reference: rgb, n, d, s, f
            nout = args.snumclass if st == 's' else HEAD_CHANNELS[st]
            if st in ('rgb', 'n', 'd'):
                setattr(self, 'deconv3' + st, _block(True, ngf * 2 * m, ngf, 4, 2, 1, bnf))
                setattr(self, 'deconv2' + st, _block(True, ngf * m, ngf // 2, 4, 2, 1, bnf))
            else:
                setattr(self, 'deconv3' + st, _block(True, ngf * 2, ngf, 4, 2, 1, bnf))
                setattr(self, 'deconv2' + st, _block(True, ngf, ngf, 4, 2, 1, bnf))
            setattr(self, 'deconv1' + st, nn.Conv2d(ngf, nout, 1, 1, 0))
        self.apply(weights_init)
        self._engine = None

    def head_channels(self):
        """[(head, channels)] in output order; the descriptor head 'f' is last (evaluation.py:137-138 computes its offset)."""
        return [(h, self.snumclass if h == 's' else HEAD_CHANNELS[h]) for h in self.heads]

    def forward(self, x):
        """x: [2P,16,H,W] float32 CUDA (NCHW, as evaluation.py:242 builds it) -> [2P, sum of head channels, H, W].
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

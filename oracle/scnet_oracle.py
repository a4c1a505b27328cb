"""TEST INFRASTRUCTURE ONLY -- plain PyTorch fp32 restatement of the reference's ``SCNet.forward``.

Oracle for SURVEY.md section 8 row M1 (``model/mymodel.py:141-380``): a functional forward driven by a
``state_dict`` with the reference's key names (``conv1rgb.0.weight``, ``conv1rgb.1.weight/bias`` ...), which also
returns every intermediate activation so the CUDA layers can be checked one by one.

Pinning: ``tests/golden/make_scnet_golden.py`` loads the reference ``nn.Module`` itself (it imports as-is on CPU),
gives both the same ``state_dict`` and checks this file against it to 1e-5 max-abs before freezing the golden
output (``tests/golden/scnet_golden.npz``).  Nothing in the product path imports this module.

Semantics that matter: every conv/deconv block is conv -> BatchNorm with *batch*
statistics (``track_running_stats=False``, mymodel.py:19,32) over the images of ONE forward call -> LeakyReLU(0.1);
the three encoder blocks are shared by the view and the warped other view but each call has its own statistics
(mymodel.py:266-288).  The reference always forwards one scan pair (2 images), so a batch of P pairs is P
independent BN groups of 2 images: ``forward`` therefore processes the batch pair by pair.
"""
import torch
import torch.nn.functional as F

BN_EPS = 1e-5        # nn.BatchNorm2d default
LEAKY = 0.1          # mymodel.py:20,33


def _block(sd, name, x, transposed, stride, padding, trace=None, tag=None):
    w = sd[name + '.0.weight']
    b = sd.get(name + '.0.bias')                     # batchnorm=0 blocks: biased convolution, no BatchNorm (mymodel.py:23-27,35-39)
    if transposed:
        y = F.conv_transpose2d(x, w, None, stride=stride, padding=padding)
    else:
        y = F.conv2d(x, w, None, stride=stride, padding=padding)
    if trace is not None:
        trace[(tag or name) + ':raw'] = y
    if b is not None:
        y = y + b.view(1, -1, 1, 1)
    else:
        y = F.batch_norm(y, None, None, sd[name + '.1.weight'], sd[name + '.1.bias'], True, 0.0, BN_EPS)
    y = F.leaky_relu(y, LEAKY)
    if trace is not None:
        trace[(tag or name) + ':act'] = y
    return y


def conv(sd, name, x, k, s, p, trace=None, tag=None):
    return _block(sd, name, x, False, s, p, trace, tag)


def deconv(sd, name, x, k, s, p, trace=None, tag=None):
    return _block(sd, name, x, True, s, p, trace, tag)


def forward_pair(sd, x, snumclass, use_tanh=True, trace=None, skip=True, heads=('rgb', 'n', 'd', 's', 'f')):
    """x: [2,16,H,W] one scan pair -> [2, sum of head channels, H, W].  mymodel.py:259-380; ``skip`` = args.skipLayer (:302 /
    :333 branches), ``heads`` = the output heads args.outputType selects, in the reference's order; a batchnorm=0 network is
    recognised by its ``.0.bias`` keys."""
    in_shape = x.shape[2:]
    x = F.interpolate(x, size=[224, 224], mode='bilinear', align_corners=False)          # :261
    if trace is not None:
        trace['in224'] = x
    own = dict(rgb=torch.cat((x[:, 0:3], x[:, 7:8]), 1), n=torch.cat((x[:, 3:6], x[:, 7:8]), 1),
               d=torch.cat((x[:, 6:7], x[:, 7:8]), 1))                                     # :264,266,270,274
    oth = dict(rgb=torch.cat((x[:, 8:11], x[:, 15:16]), 1), n=torch.cat((x[:, 11:14], x[:, 15:16]), 1),
               d=torch.cat((x[:, 14:15], x[:, 15:16]), 1))                                 # :265,278,282,286
    enc = {}
    for which, src in (('', own), ('_t2s', oth)):
        for st in ('rgb', 'n', 'd'):
            a1 = conv(sd, 'conv1' + st, src[st], 3, 1, 1, trace, 'conv1' + st + which)
            a2 = conv(sd, 'conv2' + st, a1, 4, 2, 1, trace, 'conv2' + st + which)
            a3 = conv(sd, 'conv3' + st, a2, 4, 2, 1, trace, 'conv3' + st + which)
            enc[st + which] = (a1, a2, a3)
    xin = torch.cat((enc['rgb'][2], enc['rgb_t2s'][2], enc['n'][2], enc['n_t2s'][2], enc['d'][2], enc['d_t2s'][2]), 1)  # :291
    x4 = conv(sd, 'conv4', xin, 4, 2, 1, trace)
    x5 = conv(sd, 'conv5', x4, 4, 2, 1, trace)
    x6 = conv(sd, 'conv6', x5, 4, 2, 1, trace)
    x7 = conv(sd, 'conv7', x6, 3, 2, 0, trace)
    x8 = conv(sd, 'conv8', x7, 3, 1, 1, trace)
    x9 = conv(sd, 'conv9', x8, 3, 1, 0, trace)
    cat = (lambda a, b: torch.cat((a, b), 1)) if skip else (lambda a, b: a)
    dx9 = deconv(sd, 'deconv9', x9, 3, 1, 0, trace)                                      # :302-307 / :333-339
    dx8 = deconv(sd, 'deconv8', cat(dx9, x8), 3, 1, 1, trace)
    dx7 = deconv(sd, 'deconv7', cat(dx8, x7), 3, 2, 0, trace)
    dx6 = deconv(sd, 'deconv6', cat(dx7, x6), 4, 2, 1, trace)
    dx5 = deconv(sd, 'deconv5', cat(dx6, x5), 4, 2, 1, trace)
    dx4 = deconv(sd, 'deconv4', cat(dx5, x4), 4, 2, 1, trace)
    outs = []
    for st in [h for h in ('rgb', 'n', 'd') if h in heads]:                               # :309-325 / :341-357
        e1, e2, e3 = enc[st]
        d3 = deconv(sd, 'deconv3' + st, cat(dx4, e3), 4, 2, 1, trace)
        d2 = deconv(sd, 'deconv2' + st, cat(d3, e2), 4, 2, 1, trace)
        d1 = F.conv2d(cat(d2, e1), sd['deconv1' + st + '.weight'], sd['deconv1' + st + '.bias'])
        outs.append(d1)
    for st in [h for h in ('s', 'f') if h in heads]:                                      # :364-376
        d3 = deconv(sd, 'deconv3' + st, dx4, 4, 2, 1, trace)
        d2 = deconv(sd, 'deconv2' + st, d3, 4, 2, 1, trace)
        d1 = F.conv2d(d2, sd['deconv1' + st + '.weight'], sd['deconv1' + st + '.bias'])
        if st == 'f' and use_tanh:
            d1 = torch.tanh(d1)
        outs.append(d1)
    out224 = torch.cat(outs, 1)
    if trace is not None:
        trace['out224'] = out224
    return F.interpolate(out224, size=list(in_shape), mode='bilinear', align_corners=False)   # :379


def forward(sd, x, snumclass, use_tanh=True, skip=True, heads=('rgb', 'n', 'd', 's', 'f')):
    """x: [2P,16,H,W]; consecutive image pairs are independent forward calls of the reference."""
    assert x.shape[0] % 2 == 0
    with torch.no_grad():
        return torch.cat([forward_pair(sd, x[i:i + 2], snumclass, use_tanh, None, skip, heads) for i in range(0, x.shape[0], 2)], 0)

"""Host-side sequencing of the SCNet forward (model/mymodel.py:259-380) over the C-ABI layer kernels.

torch is plumbing here: parameter storage, device buffers, the CUDA stream.  Every arithmetic op of the forward
runs in relativepose_b200/csrc/scnet.cu.  Activations are float32 NHWC "raw" conv outputs; each carries per-(pair,
channel) BN scale/shift that the consumer applies while loading (so BatchNorm + LeakyReLU + torch.cat never make
a pass over memory of their own).
"""
import ctypes

from . import _lib

NGF = 64


def h16():
    """torch dtype of the kernels' 16-bit operand / activation format (csrc/rp_h16.cuh: IEEE half by default)."""
    import torch
    return torch.float16 if _lib.h16_is_fp16() else torch.bfloat16


class _Act(object):
    """A raw NHWC activation view: buffer [2P,H,W,pitch], channel window, BN scale/shift [P,pitch]."""

    def __init__(self, buf, H, W, pitch, ch_off, C, scale=None, shift=None):
        self.buf, self.H, self.W, self.pitch, self.ch_off, self.C = buf, H, W, pitch, ch_off, C
        self.scale, self.shift = scale, shift
        self.dtype = 1 if buf.element_size() == 2 else 0          # rp_conv_src.dtype: 0 float32, 1 bfloat16

    def view(self, ch_off, C):
        return _Act(self.buf, self.H, self.W, self.pitch, self.ch_off + ch_off, C, self.scale, self.shift)


def pack_halo(w_taps, tap_widx, src_channels, cout, bn, tk):
    import torch
    return _pack_halo_f32(w_taps, tap_widx, src_channels, cout, bn, tk).to(h16()).contiguous()


def _pack_halo_f32(w_taps, tap_widx, src_channels, cout, bn, tk):
    """Weights [taps, Cin_total, Cout] fp32 -> bf16 blocks [n-tile][K chunk][block][tk/8][bn/8][8 (co)][8 (ci)] for the
    halo kernel (csrc/scnet_halo.cu): block i of a K chunk is tap ``tap_widx[i]`` (the (class, tap) order reported by
    rp_conv_halo_plan), so the blocks one CTA consumes are contiguous and in consumption order."""
    import torch
    ntn = -(-cout // bn)
    if ntn * bn != cout:                                                        # zero-pad Cout to whole n-tiles
        Wz = torch.zeros((w_taps.shape[0], w_taps.shape[1], ntn * bn), dtype=w_taps.dtype, device=w_taps.device)
        Wz[:, :, :cout] = w_taps
        w_taps = Wz
    k0s, base = [], 0
    for c in src_channels:
        assert c % tk == 0
        k0s += [base + c0 for c0 in range(0, c, tk)]
        base += c
    W = w_taps[list(tap_widx)]                                                  # [nb, Cin, Cout]
    Wt = torch.stack([W[:, k0:k0 + tk, :] for k0 in k0s], 0)                    # [kt, nb, tk, Cout]
    Wt = Wt.reshape(len(k0s), len(tap_widx), tk // 8, 8, ntn, bn // 8, 8)       # [kt, nb, kc, kk, nt, nc, r]
    return Wt.permute(4, 0, 1, 2, 5, 6, 3).contiguous()                         # [nt, kt, nb, kc, nc, r, kk]


class ScnetEngine(object):
    _act_default = 'bf16'    # storage of the BN'd activations in 'tc' mode (ResnetEngine: float32, its pooling /
                             # residual kernels are float32)

    def __init__(self, net, mode=None):
        import os
        import torch
        # 'tc': tcgen05 tensor-core kernels with 16-bit operands wherever a layer qualifies (default, the throughput mode);
        # 'tc3': the same kernels in split precision -- float32 activation storage, every tensor-core layer computed as
        #        half(x) w_hi + lo(x) w_hi + half(x) lo(w) (x w to ~2^-22, float32 accumulation): ONE launch with a hi and a lo halo
        #        and three MMAs per K step where that fits shared memory, three accumulating launches otherwise (stride-2
        #        convolutions).  float32-class output at tensor-core speed: the parity mode for the descriptors that drive the solver;
        # 'fp32': CUDA-core float32 only (the exact-parity reference path of the tests)
        self.mode = mode or os.environ.get("RP_SCNET_MODE", "tc")
        self.split3 = self.mode == 'tc3'
        # tc3: layers whose doubled halo fits shared memory take ONE fused launch, the others three (RP_SCNET_TC3=passes: always
        # three -- each term in its own accumulator: the most accurate form, 1.5e-4 instead of 2.4e-4 on the output)
        self.tc3_fused = os.environ.get("RP_SCNET_TC3", "fused") != "passes"
        if self.split3 and not _lib.h16_is_fp16():
            raise RuntimeError("RP_SCNET_MODE=tc3 needs the IEEE-half build of the library (csrc/rp_h16.cuh)")
        # replay the ~87 layer launches of a forward as one CUDA graph once a shape has been seen twice
        self.use_graph = os.environ.get("RP_SCNET_GRAPH", "1") == "1"
        # tcgen05 halo-tile kernel (csrc/scnet_halo.cu) for every layer with >= 16 channels per source in 'tc' mode
        self.halo = True
        # bit 0: 16-pixel halo pitch (debugging aid); bit 1: producer BatchNorm + LeakyReLU of 16-bit sources in packed half
        # arithmetic (default: as accurate as the float form on the reference goldens, a third of the loader instructions)
        self.halo_flags = int(os.environ.get("RP_SCNET_HALO_FLAGS", "2"))
        self.halo_min = int(os.environ.get("RP_SCNET_HALO_MIN", "1"))     # smallest base-grid extent that takes the halo kernel
        act = os.environ.get("RP_SCNET_ACT", self._act_default)
        self.act_bf16 = self.mode == 'tc' and act == 'bf16'      # (tc3 accumulates its three launches in float32 storage)
        # the forward as ONE native call: after a warm-up run the layer calls of a forward are frozen into an op list
        # (rp_net_op) that rp_scnet_forward / rp_resnet18_8s_forward replays (and that the CUDA graph captures)
        self.use_plan = os.environ.get("RP_SCNET_PLAN", "1") == "1"
        self._rec = None
        self._plans = {}
        self._graphs = {}
        self._seen = {}
        self.torch = torch
        self.net = net
        self.lib = _lib.load()
        self._packed_key = None
        self._packed = {}
        self._packed_tc = {}
        self._bufs = {}
        self._P = 0
        self._dev = None

    _forward_entry = "rp_scnet_forward"

    def _run(self, name, *args):
        """One layer call of the C ABI (last argument: the stream); recorded into the op list while a plan is being built."""
        cargs = [ctypes.byref(a) if isinstance(a, _lib.RpConvDesc) else a for a in args]
        _lib.check(getattr(self.lib, name)(*cargs), name)
        if self._rec is not None:
            op = _lib.RpNetOp()
            op.kind = _lib.NET_OPS[name]
            rest = list(args[:-1])
            if rest and isinstance(rest[0], _lib.RpConvDesc):
                op.conv = _lib.RpConvDesc.from_buffer_copy(rest[0])
                rest = rest[1:]
            for i, a in enumerate(rest):
                op.arg[i] = (0 if a is None else int(a)) & 0xFFFFFFFFFFFFFFFF
            self._rec.append(op)

    def _run_plan(self, plan):
        ops, n = plan
        _lib.check(getattr(self.lib, self._forward_entry)(ops, n, self.torch.cuda.current_stream().cuda_stream), self._forward_entry)

    # ---------------------------------------------------------------- weights
    def _pack(self):
        torch = self.torch
        key = tuple((p.data_ptr(), p._version) for p in self.net.parameters())
        if key == self._packed_key:
            return
        W = {}
        sd = dict(self.net.named_parameters())
        for name, p in sd.items():
            if name.endswith('.0.weight'):
                base = name[:-9]
                mod = getattr(self.net, base)[0]
                tr = isinstance(mod, torch.nn.ConvTranspose2d)
                w = p.detach()
                w = w.permute(2, 3, 0, 1) if tr else w.permute(2, 3, 1, 0)       # -> [k,k,Cin,Cout]
                W[base] = w.contiguous().float()
            elif name.startswith('deconv1') and name.endswith('.weight'):
                W[name[:-7]] = p.detach().permute(2, 3, 1, 0).contiguous().float()   # 1x1 heads
        self._packed = W
        self._packed_tc = {}
        self._packed_key = key

    # ---------------------------------------------------------------- buffers
    def _alloc(self, P, device, heads):
        """heads: [(name, channels)] in output order (SCNet.head_channels())."""
        torch = self.torch
        cout_total = sum(c for _, c in heads)
        if self._P == P and self._dev == device and self._bufs.get('heads') == tuple(heads):
            return
        n = 2 * P
        self._plans, self._graphs, self._seen = {}, {}, {}          # they hold pointers into the buffers replaced below
        f = dict(dtype=torch.float32, device=device)
        fa = dict(dtype=h16() if self.act_bf16 else torch.float32, device=device)
        B = {}

        def act(name, H, W, C):
            # the bottleneck tensors (<= 14x14) stay float32: they are tiny, and conv9's two-sample BatchNorm amplifies
            # any storage rounding by up to 1/sqrt(eps) (tests/test_gpu_scnet.py)
            B[name] = _Act(torch.empty((n, H, W, C), **(fa if H > 14 else f)), H, W, C, 0, C,
                           torch.empty((P, C), **f), torch.empty((P, C), **f))

        B['in20'] = _Act(torch.empty((n, 224, 224, 20), **f), 224, 224, 20, 0, 20)
        if self.act_bf16 or self.split3:
            B['in96'] = _Act(torch.empty((n, 224, 224, 96), dtype=h16(), device=device), 224, 224, 96, 0, 96)
        for st in ('rgb', 'n', 'd'):
            for wh in ('', '_t2s'):
                act('e1' + st + wh, 224, 224, 32)
                act('e2' + st + wh, 112, 112, 64)
        act('xin', 56, 56, 768)
        act('x4', 28, 28, 256); act('x5', 14, 14, 512); act('x6', 7, 7, 512)
        act('x7', 3, 3, 512); act('x8', 3, 3, 512); act('x9', 1, 1, 1024)
        act('dx9', 3, 3, 512); act('dx8', 3, 3, 512); act('dx7', 7, 7, 512)
        act('dx6', 14, 14, 512); act('dx5', 28, 28, 256); act('dx4', 56, 56, 128)
        for st, _ in heads:
            act('d3' + st, 112, 112, 64)
            act('d2' + st, 224, 224, 32 if st in ('rgb', 'n', 'd') else 64)
        # 224x224 output of the heads, each at a 16-byte aligned channel offset (float4 stores in the head kernels)
        offs, cmap, o = {}, [], 0
        for st, c in heads:
            offs[st] = o
            cmap += list(range(o, o + c))
            o += 4 * ((c + 3) // 4)
        pitch = o
        B['out224'] = _Act(torch.zeros((n, 224, 224, pitch), **f), 224, 224, pitch, 0, pitch)
        B['head_off'] = offs
        B['cmap'] = torch.tensor(cmap, dtype=torch.int32, device=device)
        B['ctot'] = cout_total
        B['heads'] = tuple(heads)
        B['partials'] = None
        self._bufs, self._P, self._dev = B, P, device

    # ---------------------------------------------------------------- one conv block
    _slope = 0.1          # LeakyReLU(0.1) of the SCNet blocks; ResnetEngine overrides with 0 (ReLU)
    _gsz = 2              # images per BatchNorm batch (one scan pair)

    def _conv(self, name, srcs, out, transposed, k, s, p, bn=True, bias=None, tanh=False, stream=None, bn_params=None, wkey=None,
              block_bias=None):
        """block_bias: the layer is a batchnorm=0 block (mymodel.py:23-27,35-39: biased convolution + LeakyReLU).  The raw
        convolution is stored and the consumer's load-time transform gets scale = 1, shift = bias -- the same mechanism that
        applies a BatchNorm, so no kernel knows the difference."""
        torch = self.torch
        if block_bias is not None:
            bn = False
            sl = slice(out.ch_off, out.ch_off + out.C)
            out.scale[:, sl] = 1.0                       # (eager runs only; a frozen plan / graph replays the kernels, the
            out.shift[:, sl] = block_bias.detach().float()   #  two small tensors keep their values)
        wkey = wkey or name                  # entry of self._packed holding this layer's [k,k,Cin,Cout] weights
        d = _lib.RpConvDesc()
        d.imgs_per_group = self._gsz
        d.nsrc = len(srcs)
        for i, a in enumerate(srcs):
            d.src[i].ptr = a.buf.data_ptr()
            d.src[i].pitch, d.src[i].ch_off, d.src[i].C = a.pitch, a.ch_off, a.C
            d.src[i].dtype = a.dtype
            if a.scale is not None:
                d.src[i].act = 1
                d.src[i].slope = self._slope
                d.src[i].scale, d.src[i].shift = a.scale.data_ptr(), a.shift.data_ptr()
                d.src[i].sstride, d.src[i].s_off = a.pitch, a.ch_off
            else:
                d.src[i].act = 0
        d.transposed, d.k, d.s, d.p = int(transposed), k, s, p
        d.G = self._P
        d.Hin, d.Win, d.Hout, d.Wout = srcs[0].H, srcs[0].W, out.H, out.W
        d.Cout = out.C
        d.W = self._packed[wkey].data_ptr()
        d.out, d.out_pitch, d.out_ch_off = out.buf.data_ptr(), out.pitch, out.ch_off
        d.out_dtype = out.dtype
        d.bias = bias.data_ptr() if bias is not None else None
        d.tanh_out = int(tanh)
        use_tc = self.mode in ('tc', 'tc3') and all(a.C % 16 == 0 for a in srcs)
        if k == 1 and not bn and sum(a.C for a in srcs) <= 128 and (out.C <= 4 or (out.C <= 32 and not self.halo)):
            use_tc = False                   # 3-channel heads are HBM-bound: CUDA-core kernel inside rp_conv_layer (also the
                                             # fallback for the wider heads when the halo kernel is off)
        use_halo = False
        nparts = ctypes.c_int(0)
        if use_tc and self.halo and (k in (3, 4) or (k == 1 and s in (1, 2) and not transposed)) and \
                min(out.H, out.W) // (s if transposed else 1) >= self.halo_min:
            # halo-tile kernel: stride-2 convolutions keep 4 parity planes of the halo, so their K chunk is 32
            tk = 32 if (s == 2 and not transposed) else (64 if all(a.C % 64 == 0 for a in srcs) else 32)
            if any(a.C % 32 for a in srcs):
                tk = 16                          # the bf16-split stem (conv1*, 16 channels per group)
            cap = 64 if (transposed and s == 2) else 128                  # 4 accumulators x bn TMEM columns
            if s == 2 and not transposed:
                cap = 256                                                 # one accumulator set of 256 columns (x2 sets = all of TMEM)
            bn_tile = next((b for b in (256, 128, 64, 32) if b <= cap and out.C % b == 0), 32)   # 32: Cout zero-padded (heads)
            ntap = ctypes.c_int(0)
            widx = (ctypes.c_int * 16)()
            split3 = self.split3 and not wkey.endswith('#split')      # the stem input is already a hi/lo split: one plain launch
            fused_ok = False
            if split3 and self.tc3_fused:
                # the fused split-precision launch keeps a hi and a lo halo per buffer: take the largest K chunk whose doubled
                # halo still fits next to the weight ring (else the layer runs as three launches at its usual K chunk)
                for tk_try in ([64, 32] if tk == 64 else [tk]):
                    if self.lib.rp_conv_halo_fits(ctypes.byref(d), bn_tile, tk_try, self.halo_flags | (1 << 10)) == _lib.RP_OK:
                        tk, fused_ok = tk_try, True
                        break
            if bn_tile and self.lib.rp_conv_halo_plan(ctypes.byref(d), bn_tile, tk, self.halo_flags, ctypes.byref(nparts),
                                                      ctypes.byref(ntap), widx) == _lib.RP_OK:
                use_halo = True
                key = (wkey, 'halo', bn_tile, tk)
                if key not in self._packed_tc:
                    w = self._packed[wkey]
                    self._packed_tc[key] = pack_halo(w.reshape(k * k, w.shape[2], w.shape[3]), list(widx[:ntap.value]),
                                                     [a.C for a in srcs], out.C, bn_tile, tk)
                wtc = self._packed_tc[key]
                fused3 = fused_ok
                if split3:
                    # split precision: (hi, lo) block pairs of 2^8 w (both halves in the normal range of IEEE half) for the fused
                    # launch / the first of two launches, the hi blocks alone for the second
                    key_f = key + ('f3',)
                    if key_f not in self._packed_tc:
                        w = self._packed[wkey].reshape(k * k, self._packed[wkey].shape[2], self._packed[wkey].shape[3]) * 256.0
                        w_hi = w.to(h16()).float()
                        taps = list(widx[:ntap.value])
                        both = torch.cat((w_hi, w - w_hi), 0)                                       # [2 k k, Cin, Cout]
                        order = [i for t in taps for i in (t, k * k + t)]
                        self._packed_tc[key_f] = pack_halo(both, order, [a.C for a in srcs], out.C, bn_tile, tk)
                        self._packed_tc[key + ('hi8',)] = pack_halo(w_hi, taps, [a.C for a in srcs], out.C, bn_tile, tk)
                        self._packed_tc[key + ('lo8',)] = pack_halo(w - w_hi, taps, [a.C for a in srcs], out.C, bn_tile, tk)
                    wtc_plain = wtc
                    wtc, wtc_hi, wtc_lo = self._packed_tc[key_f], self._packed_tc[key + ('hi8',)], self._packed_tc[key + ('lo8',)]
        if use_tc and not use_halo:
            use_tc = False                  # no tile plan for this shape: the float32 CUDA-core kernel (reads / writes 16-bit storage too)
        if bn:
            if not use_halo:
                _lib.check(self.lib.rp_conv_nparts(ctypes.byref(d), ctypes.byref(nparts)), "rp_conv_nparts")
            need = self._P * nparts.value * out.C
            pt = self._bufs['partials']
            if pt is None or pt.numel() < 2 * need:
                pt = torch.empty((2 * need,), dtype=torch.float32, device=self._dev)
                self._bufs['partials'] = pt
            d.psum, d.psq = pt.data_ptr(), pt.data_ptr() + 4 * need
        else:
            d.psum, d.psq = None, None
        if use_halo and fused3:
            self._run("rp_conv_layer_halo", d, wtc.data_ptr(), bn_tile, tk, self.halo_flags | (1 << 10), stream)
        elif use_halo and split3:
            # the doubled halo does not fit shared memory (stride-2 convolutions): three launches, each term in an accumulator of
            # its own -- half(x) hi(w) stores (the plain 16-bit launch over float32 storage), lo(x') hi(w') (flags bits 8 + 9) and
            # half(x') lo(w') (bit 9) add 2^-12 x their accumulators onto the float32 output; bias / tanh / batch statistics belong to
            # the last launch, which sees the complete sum.  (Measured: more accurate than two launches with weight block pairs,
            # flags bit 11 -- the tensor core's float32 accumulator loses the low-order terms it is handed one by one.)
            d0 = _lib.RpConvDesc.from_buffer_copy(d)
            d0.bias, d0.tanh_out, d0.psum, d0.psq = None, 0, None, None
            self._run("rp_conv_layer_halo", d0, wtc_plain.data_ptr(), bn_tile, tk, self.halo_flags, stream)
            self._run("rp_conv_layer_halo", d0, wtc_hi.data_ptr(), bn_tile, tk, self.halo_flags | (1 << 8) | (1 << 9), stream)
            self._run("rp_conv_layer_halo", d, wtc_lo.data_ptr(), bn_tile, tk, self.halo_flags | (1 << 9), stream)
        elif use_halo:
            self._run("rp_conv_layer_halo", d, wtc.data_ptr(), bn_tile, tk, self.halo_flags, stream)
        else:
            self._run("rp_conv_layer", d, stream)
        if bn:
            if bn_params is None:
                bnm = getattr(self.net, name)[1]
                bn_params = (bnm.weight, bnm.bias)
            if nparts.value >= 4096:              # one long list (Resnet18_8s): slice it over more blocks
                nsplit = 32
                sc = self._bufs.get('bn_scratch')
                if sc is None or sc.numel() < self._P * nsplit * out.C * 2:
                    sc = torch.empty((self._P * nsplit * out.C * 2,), dtype=torch.float64, device=self._dev)
                    self._bufs['bn_scratch'] = sc
                self._run("rp_bn_finalize_split", d.psum, d.psq, self._P, nparts.value, out.C, self._gsz * out.H * out.W,
                                                         bn_params[0].data_ptr(), bn_params[1].data_ptr(), out.scale.data_ptr(),
                                                         out.shift.data_ptr(), out.pitch, out.ch_off, nsplit, sc.data_ptr(), stream)
            else:
                self._run("rp_bn_finalize", d.psum, d.psq, self._P, nparts.value, out.C, self._gsz * out.H * out.W,
                                                   bn_params[0].data_ptr(), bn_params[1].data_ptr(),
                                                   out.scale.data_ptr(), out.shift.data_ptr(), out.pitch, out.ch_off, stream)

    # ---------------------------------------------------------------- forward
    def forward(self, x, trace=None, borrow=False, heads=None):
        """``borrow=True`` returns the engine's own output buffer (valid until the next forward of this engine) instead of a
        fresh copy -- the batched pipeline consumes the output before it calls the network again, and the copy of a
        [64,54,160,640] float32 tensor is 2.8 GB of HBM traffic per call.  ``heads``: compute only these output heads (their
        decoder branches, 1x1 heads and the resize of their channels); the channels of the other heads in the returned tensor
        are then undefined -- the alternation reads normals, depth and descriptors only, the rgb and semantic branches are a
        fifth of the forward."""
        torch = self.torch
        heads = tuple(heads) if heads else None
        if heads is not None and set(heads) >= set(h for h, _ in self.net.head_channels()):
            heads = None
        if (self.use_plan or self.use_graph) and trace is None and x.is_cuda and x.dim() == 4:
            key = (tuple(x.shape), str(x.device), self.mode, heads, tuple((p.data_ptr(), p._version) for p in self.net.parameters()))
            ent = self._graphs.get(key)
            if ent is not None:                       # CUDA-graph replay of the native forward
                gph, xs, ys = ent
                if x.data_ptr() != xs.data_ptr():
                    xs.copy_(x)
                gph.replay()
                return ys if borrow else ys.clone()
            ent = self._plans.get(key)
            if ent is not None:
                xs, ys, plan = ent
                if self.use_graph:                    # third call: capture the one native call into a graph
                    with torch.cuda.device(x.device):
                        torch.cuda.synchronize()
                        gph = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(gph):
                            self._run_plan(plan)
                    self._graphs = {key: (gph, xs, ys)}
                    if x.data_ptr() != xs.data_ptr():
                        xs.copy_(x)
                    gph.replay()
                    return ys if borrow else ys.clone()
                with torch.cuda.device(x.device):
                    if x.data_ptr() != xs.data_ptr():
                        xs.copy_(x)
                    self._run_plan(plan)
                return ys if borrow else ys.clone()
            self._seen[key] = self._seen.get(key, 0) + 1
            if self._seen[key] == 2:                  # the first run sized every buffer: freeze the second one into a plan
                xs = x.contiguous().float().clone()
                self._rec = []
                try:
                    ys = self._forward_eager(xs, None, heads)
                    rec = self._rec
                finally:
                    self._rec = None
                self._plans = {key: (xs, ys, ((_lib.RpNetOp * len(rec))(*rec), len(rec)))}     # one plan: buffers are shared between shapes
                return ys if borrow else ys.clone()
        return self._forward_eager(x, trace, heads)

    def input_buffer(self, shape, device):
        """The static input tensor of the frozen plan for this shape (None before the plan exists): a caller that assembles
        its network input in place (pipeline.py) writes here and skips the input copy as well."""
        for key, ent in list(self._graphs.items()) + list(self._plans.items()):
            if key[0] == tuple(shape) and key[1] == str(device):
                return ent[1] if key in self._graphs else ent[0]
        return None

    def _forward_eager(self, x, trace=None, only=None):
        torch = self.torch
        if not x.is_cuda:
            raise RuntimeError("relativepose_b200.SCNet.forward needs a CUDA tensor (no CPU fallback)")
        if x.dim() != 4 or x.shape[1] != 16 or x.shape[0] % 2 != 0:
            raise ValueError("expected [2P,16,H,W]")
        x = x.contiguous().float()
        n, _, H, W = x.shape
        P = n // 2
        net = self.net
        heads = net.head_channels()                      # [(name, channels)] of the heads args.outputType selected, output order
        ctot = sum(c for _, c in heads)
        skip = bool(net.skipLayer)
        nobn = not bool(getattr(net, 'batchnorm', 1))    # batchnorm=0: biased convolutions, no BatchNorm (mymodel.py:23-27,35-39)
        with torch.cuda.device(x.device), torch.no_grad():
            self._pack()
            self._alloc(P, x.device, heads)
            B = self._bufs
            stream = torch.cuda.current_stream().cuda_stream

            def block(name, srcs, out, transposed, k, s, p, **kw):
                """One conv2d / deconv2d block of the reference (mymodel.py:15-39)."""
                if nobn:
                    kw['block_bias'] = getattr(net, name)[0].bias
                self._conv(name, srcs, out, transposed, k, s, p, stream=stream, **kw)

            def cat(a, b):
                return [a, b] if skip else [a]           # skipLayer=0: the decoder sees no encoder tensors (:333-357)
            split_stem = (self.act_bf16 or self.split3) and self.halo     # conv1* on tcgen05 from the 16-bit hi/lo split input
            if split_stem:
                self._run("rp_scnet_resize_in_split", x.data_ptr(), n, H, W, B['in96'].buf.data_ptr(), stream)
                for st in ('rgb', 'n', 'd'):
                    if 'conv1' + st + '#split' not in self._packed:
                        w = self._packed['conv1' + st]                          # [3,3,cin,32]
                        hi = w.to(h16()).float()
                        lo = (w - hi).to(h16()).float()
                        ws = torch.zeros((3, 3, 16, 32), dtype=torch.float32, device=w.device)
                        c = w.shape[2]
                        ws[:, :, 0:c], ws[:, :, 4:4 + c], ws[:, :, 8:8 + c] = hi, hi, lo   # x [hi|lo|hi|0] . w [hi|hi|lo|0]
                        self._packed['conv1' + st + '#split'] = ws.contiguous()
            else:
                self._run("rp_scnet_resize_in", x.data_ptr(), n, H, W, B['in20'].buf.data_ptr(), stream)
            chan = {'rgb': (0, 4), 'n': (4, 4), 'd': (8, 2)}
            xin_slot = {'rgb': 0, 'rgb_t2s': 128, 'n': 256, 'n_t2s': 384, 'd': 512, 'd_t2s': 640}
            for wh, base in (('', 0), ('_t2s', 10)):
                for st in ('rgb', 'n', 'd'):
                    off, c = chan[st]
                    if split_stem:
                        grp = {'rgb': 0, 'n': 1, 'd': 2}[st] + (3 if wh else 0)
                        block('conv1' + st, [B['in96'].view(16 * grp, 16)], B['e1' + st + wh], False, 3, 1, 1, wkey='conv1' + st + '#split')
                    else:
                        block('conv1' + st, [B['in20'].view(base + off, c)], B['e1' + st + wh], False, 3, 1, 1)
                    block('conv2' + st, [B['e1' + st + wh]], B['e2' + st + wh], False, 4, 2, 1)
                    block('conv3' + st, [B['e2' + st + wh]], B['xin'].view(xin_slot[st + wh], 128), False, 4, 2, 1)
            block('conv4', [B['xin']], B['x4'], False, 4, 2, 1)
            block('conv5', [B['x4']], B['x5'], False, 4, 2, 1)
            block('conv6', [B['x5']], B['x6'], False, 4, 2, 1)
            block('conv7', [B['x6']], B['x7'], False, 3, 2, 0)
            block('conv8', [B['x7']], B['x8'], False, 3, 1, 1)
            block('conv9', [B['x8']], B['x9'], False, 3, 1, 0)
            block('deconv9', [B['x9']], B['dx9'], True, 3, 1, 0)
            block('deconv8', cat(B['dx9'], B['x8']), B['dx8'], True, 3, 1, 1)
            block('deconv7', cat(B['dx8'], B['x7']), B['dx7'], True, 3, 2, 0)
            block('deconv6', cat(B['dx7'], B['x6']), B['dx6'], True, 4, 2, 1)
            block('deconv5', cat(B['dx6'], B['x5']), B['dx5'], True, 4, 2, 1)
            block('deconv4', cat(B['dx5'], B['x4']), B['dx4'], True, 4, 2, 1)
            for st, c in heads:
                if only is not None and st not in only:
                    continue                                                   # this head's decoder branch is not wanted
                o = B['head_off'][st]
                if st in ('rgb', 'n', 'd'):                                    # mymodel.py:309-325 (skipLayer=1 only, see SCNet.__init__)
                    block('deconv3' + st, [B['dx4'], B['xin'].view(xin_slot[st], 128)], B['d3' + st], True, 4, 2, 1)
                    block('deconv2' + st, [B['d3' + st], B['e2' + st]], B['d2' + st], True, 4, 2, 1)
                    self._conv('deconv1' + st, [B['d2' + st], B['e1' + st]], B['out224'].view(o, c), False, 1, 1, 0, bn=False,
                               bias=getattr(net, 'deconv1' + st).bias, stream=stream)
                else:                                                          # :364-376
                    block('deconv3' + st, [B['dx4']], B['d3' + st], True, 4, 2, 1)
                    block('deconv2' + st, [B['d3' + st]], B['d2' + st], True, 4, 2, 1)
                    self._conv('deconv1' + st, [B['d2' + st]], B['out224'].view(o, c), False, 1, 1, 0, bn=False,
                               bias=getattr(net, 'deconv1' + st).bias, tanh=(st == 'f' and bool(net.useTanh)), stream=stream)
            out = torch.empty((n, ctot, H, W), dtype=torch.float32, device=x.device)
            if only is None:
                self._run("rp_scnet_resize_out_map", B['out224'].buf.data_ptr(), n, B['out224'].pitch, B['cmap'].data_ptr(), ctot, H, W,
                                                            out.data_ptr(), stream)
            else:
                c0, items = 0, []                          # (source channel of out224, output channel) of every wanted channel
                cm = B['cmap'].tolist() if 'cmap_list' not in B else B['cmap_list']
                B['cmap_list'] = cm
                for st, c in heads:
                    if st in only:
                        items += [(cm[c0 + j], c0 + j) for j in range(c)]
                    c0 += c
                skey = ('submap', tuple(only))
                if skey not in B:
                    B[skey] = (torch.tensor([a for a, _ in items], dtype=torch.int32, device=x.device),
                               torch.tensor([b for _, b in items], dtype=torch.int32, device=x.device))
                sm, om = B[skey]
                self._run("rp_scnet_resize_out_sub", B['out224'].buf.data_ptr(), n, B['out224'].pitch, sm.data_ptr(), om.data_ptr(), len(items),
                                                            H, W, out.data_ptr(), ctot, stream)
            if trace is not None:
                self._dump(trace)
        return out

    def _dump(self, trace):
        """Test hook: every intermediate as NCHW tensors (raw and BN+LeakyReLU'd), named like the oracle's trace."""
        torch = self.torch
        B = self._bufs

        def grab(a):
            raw = a.buf[..., a.ch_off:a.ch_off + a.C].float()
            out = {'raw': raw.permute(0, 3, 1, 2).contiguous()}
            if a.scale is not None:
                sc = a.scale[:, a.ch_off:a.ch_off + a.C].repeat_interleave(2, 0)[:, None, None, :]
                sh = a.shift[:, a.ch_off:a.ch_off + a.C].repeat_interleave(2, 0)[:, None, None, :]
                out['act'] = torch.nn.functional.leaky_relu(raw * sc + sh, 0.1).permute(0, 3, 1, 2).contiguous()
            return out

        slot = {'rgb': 0, 'rgb_t2s': 128, 'n': 256, 'n_t2s': 384, 'd': 512, 'd_t2s': 640}
        names = {}
        for st in ('rgb', 'n', 'd'):
            for wh in ('', '_t2s'):
                names['conv1' + st + wh] = B['e1' + st + wh]
                names['conv2' + st + wh] = B['e2' + st + wh]
                names['conv3' + st + wh] = B['xin'].view(slot[st + wh], 128)
        for a, b in (('conv4', 'x4'), ('conv5', 'x5'), ('conv6', 'x6'), ('conv7', 'x7'), ('conv8', 'x8'), ('conv9', 'x9'),
                     ('deconv9', 'dx9'), ('deconv8', 'dx8'), ('deconv7', 'dx7'), ('deconv6', 'dx6'), ('deconv5', 'dx5'),
                     ('deconv4', 'dx4')):
            names[a] = B[b]
        for st, _ in B['heads']:
            names['deconv3' + st] = B['d3' + st]
            names['deconv2' + st] = B['d2' + st]
        for k, a in names.items():
            g = grab(a)
            trace[k + ':raw'] = g['raw']
            trace[k + ':act'] = g['act']
        trace['out224'] = B['out224'].buf[..., B['cmap'].long()].permute(0, 3, 1, 2).contiguous()
        trace['in20'] = B['in20'].buf.permute(0, 3, 1, 2).contiguous()

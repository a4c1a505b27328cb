"""CPU check of the halo-tile convolution plan (csrc/scnet_halo.cu): the host-side tile plan (parity planes, per-tap
descriptor offsets, weight-block order) and scnet_engine.pack_halo are replayed in numpy with exactly the address
arithmetic the kernel and the UMMA shared-memory descriptors use (core matrix = 8 rows x 16 bytes, SBO = halo row pitch,
LBO = K-core stride), and the result must equal torch's conv2d / conv_transpose2d.  No GPU needed: only the C-ABI plan
query runs natively."""
import ctypes

import numpy as np
import pytest

TH, TW = 16, 8


def _plan(d, bn, tk, flags=0):
    from relativepose_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_int * 96)()
    rc = lib.rp_conv_halo_debug(ctypes.byref(d), bn, tk, flags, out)
    if rc != 0:
        return None
    o = list(out)
    keys = ("istr", "ostr", "nclass", "nty", "ntx", "tiles_m", "nplane", "PH", "PW", "NPX", "a_lbo", "ntap", "nkt", "nkt0", "tmem_cols")
    P = dict(zip(keys, o[:15]))
    P["plane"] = [o[16 + 4 * i:20 + 4 * i] for i in range(4)]
    P["tap"] = [o[32 + 3 * i:35 + 3 * i] for i in range(16)]
    P["cls"] = [o[80 + 4 * i:84 + 4 * i] for i in range(4)]
    nparts, ntap = ctypes.c_int(0), ctypes.c_int(0)
    widx = (ctypes.c_int * 16)()
    assert lib.rp_conv_halo_plan(ctypes.byref(d), bn, tk, flags, ctypes.byref(nparts), ctypes.byref(ntap), widx) == 0
    P["widx"] = list(widx[:ntap.value])
    P["nparts"] = nparts.value
    return P


def _emulate(x, Wp, P, bn, tk, Cout, gsz, G, Hin, Win, Hout, Wout):
    """x [n,Hin,Win,Cin] float32 (already activated, sources concatenated); Wp [ntn,nkt,ntap,tk/8,bn/8,8,8] float32."""
    KC = tk // 8
    units = P["a_lbo"] // 16
    out = np.full((G * gsz, Hout, Wout, Cout), np.nan, np.float32)
    ppl = P["PH"] * P["PW"]
    m = np.arange(128)
    for g in range(G):
        for tile_m in range(P["tiles_m"]):
            im, trem = divmod(tile_m, P["nty"] * P["ntx"])
            tyi, txi = divmod(trem, P["ntx"])
            a0, b0 = tyi * TH, txi * TW
            img = g * gsz + im
            # halo pixel table
            pix = np.full(P["NPX"], -1, np.int64)
            for h in range(P["NPX"]):
                p, r = divmod(h, ppl)
                hy, hx = divmod(r, P["PW"])
                qy, qx, oy, ox = P["plane"][p]
                iy = (a0 + oy + hy) * P["istr"] + qy
                ix = (b0 + ox + hx) * P["istr"] + qx
                if 0 <= iy < Hin and 0 <= ix < Win:
                    pix[h] = iy * Win + ix
            for tile_n in range(-(-Cout // bn)):
                acc = np.zeros((P["nclass"], 128, bn), np.float64)
                seen = np.zeros(P["nclass"], bool)
                for c in range(P["nkt"]):
                    halo = np.full((KC, units, 8), np.nan, np.float32)          # NaN = never written: must never be read
                    xi = x[img].reshape(Hin * Win, -1)
                    for kc in range(KC):
                        ch = c * tk + kc * 8
                        v = np.zeros((P["NPX"], 8), np.float32)
                        ok = pix >= 0
                        v[ok] = xi[pix[ok], ch:ch + 8]
                        halo[kc, :P["NPX"]] = v
                    for t in range(P["ntap"]):
                        cls, a_off, first = P["tap"][t]
                        blk = Wp[tile_n, c, t]                                    # [KC, bn/8, 8 (co), 8 (ci)]
                        Bm = blk.transpose(1, 2, 0, 3).reshape(bn, tk)            # B[n, k]
                        unit = a_off // 16 + (m // 8) * P["PW"] + (m % 8)
                        Am = halo[:, unit, :].transpose(1, 0, 2).reshape(128, tk)   # A[m, k]
                        if c == 0 and first:
                            assert not seen[cls]
                            seen[cls] = True
                            acc[cls] = 0.0
                        acc[cls] += Am.astype(np.float64) @ Bm.astype(np.float64).T
                assert seen.all()
                a, b = a0 + m // 8, b0 + m % 8
                for cls in range(P["nclass"]):
                    py, px, Ha, Wb = P["cls"][cls]
                    ok = (a < Ha) & (b < Wb)
                    oy, ox = a[ok] * P["ostr"] + py, b[ok] * P["ostr"] + px
                    nco = min(bn, Cout - tile_n * bn)
                    out[img, oy, ox, tile_n * bn:tile_n * bn + nco] = acc[cls][ok][:, :nco]
    return out


CASES = [
    # name, transposed, k, s, p, Hin, Win, [Cin...], Cout, bn, tk
    ("conv3x3s1", 0, 3, 1, 1, 20, 19, [64], 64, 64, 64),
    ("conv4x4s2", 0, 4, 2, 1, 36, 20, [32], 64, 32, 32),
    ("conv3x3s2p1", 0, 3, 2, 1, 33, 18, [64], 32, 32, 32),
    ("conv3x3s2p0", 0, 3, 2, 0, 35, 21, [32], 32, 32, 32),
    ("deconv4x4s2_cat", 1, 4, 2, 1, 18, 9, [64, 64], 64, 64, 64),
    ("deconv4x4s2_tk32", 1, 4, 2, 1, 17, 10, [32, 64], 32, 32, 32),
    ("deconv3x3s1", 1, 3, 1, 1, 17, 9, [64], 128, 128, 64),
    ("deconv3x3s2p0", 1, 3, 2, 0, 16, 9, [64], 32, 32, 64),
    ("conv3x3s1_pitch16", 0, 3, 1, 1, 20, 19, [64], 64, 64, 64),
    ("conv3x3s1_tk16_stem", 0, 3, 1, 1, 20, 19, [16], 32, 32, 16),
    ("head1x1_cout3", 0, 1, 1, 0, 20, 19, [32, 32], 3, 32, 32),
    ("head1x1_cout21", 0, 1, 1, 0, 17, 9, [64], 21, 32, 64),
    ("conv3x3p0_to1x1", 0, 3, 1, 0, 3, 3, [64], 128, 128, 64),
    ("deconv3x3p0_from1x1", 1, 3, 1, 0, 1, 1, [64], 64, 64, 64),
    ("conv4x4s2_14to7", 0, 4, 2, 1, 14, 14, [32], 32, 32, 32),
    ("deconv3x3s2p0_3to7", 1, 3, 2, 0, 3, 3, [64, 64], 64, 64, 64),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_halo_plan_reproduces_convolution(case):
    import torch
    import torch.nn.functional as F
    from relativepose_b200 import _lib
    from relativepose_b200.scnet_engine import pack_halo
    name, tr, k, s, p, Hin, Win, cins, Cout, bn, tk = case
    flags = 1 if name.endswith("pitch16") else 0
    G, gsz = 2, 2
    Cin = sum(cins)
    if tr:
        Hout, Wout = (Hin - 1) * s - 2 * p + k, (Win - 1) * s - 2 * p + k
    else:
        Hout, Wout = (Hin + 2 * p - k) // s + 1, (Win + 2 * p - k) // s + 1
    d = _lib.RpConvDesc()
    d.nsrc = len(cins)
    for i, c in enumerate(cins):
        d.src[i].C, d.src[i].pitch, d.src[i].ch_off = c, c, 0
    d.transposed, d.k, d.s, d.p, d.G, d.imgs_per_group = tr, k, s, p, G, gsz
    d.Hin, d.Win, d.Hout, d.Wout, d.Cout, d.out_pitch = Hin, Win, Hout, Wout, Cout, Cout
    P = _plan(d, bn, tk, flags)
    assert P is not None
    assert P["nparts"] == P["nclass"] * P["tiles_m"] and P["ntap"] <= 16 and P["NPX"] <= 640
    assert (P["a_lbo"] // 16) % 8 == 1 and P["a_lbo"] // 16 >= P["NPX"]
    if flags:
        assert P["PW"] == 16
    rng = np.random.RandomState(1)
    x = rng.randn(G * gsz, Cin, Hin, Win).astype(np.float32)
    if tr:
        w = rng.randn(Cin, Cout, k, k).astype(np.float32)
        ref = F.conv_transpose2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), stride=s, padding=p)
        wt = torch.from_numpy(w).permute(2, 3, 0, 1)                 # [k,k,Cin,Cout] as ScnetEngine._pack
    else:
        w = rng.randn(Cout, Cin, k, k).astype(np.float32)
        ref = F.conv2d(torch.from_numpy(x).double(), torch.from_numpy(w).double(), stride=s, padding=p)
        wt = torch.from_numpy(w).permute(2, 3, 1, 0)
    assert tuple(ref.shape[2:]) == (Hout, Wout)
    Wp = pack_halo(wt.reshape(k * k, Cin, Cout).contiguous().float(), P["widx"], cins, Cout, bn, tk)
    assert Wp.shape == (-(-Cout // bn), P["nkt"], P["ntap"], tk // 8, bn // 8, 8, 8)
    # the kernel reads bf16 weights; emulate with the unrounded values (this test is about addressing, not rounding)
    import relativepose_b200.scnet_engine as se
    Wf = se._pack_halo_f32(wt.reshape(k * k, Cin, Cout).contiguous().float(), P["widx"], cins, Cout, bn, tk)
    got = _emulate(x.transpose(0, 2, 3, 1).copy(), Wf.numpy(), P, bn, tk, Cout, gsz, G, Hin, Win, Hout, Wout)
    assert not np.isnan(got).any(), "output pixels never written / halo units read before written"
    err = np.abs(got - ref.permute(0, 2, 3, 1).numpy()).max()
    assert err <= 1e-3, err

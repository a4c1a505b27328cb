"""Run single SCNet-sized layers on the halo kernel (G scan pairs) for ncu: stem, conv2, conv4, deconv3, deconv2, head."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relativepose_b200.scnet_engine import ScnetEngine, _Act, h16
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda:0")
LAYERS = [
    # name, transposed, k, s, p, Hin, [Cin...], Cout, bn
    ("stem", 0, 3, 1, 1, 224, [16], 32, True),
    ("conv2", 0, 4, 2, 1, 224, [32], 64, True),
    ("conv3", 0, 4, 2, 1, 112, [64], 128, True),
    ("conv4", 0, 4, 2, 1, 56, [768], 256, True),
    ("deconv4", 1, 4, 2, 1, 28, [256, 256], 128, True),
    ("deconv3", 1, 4, 2, 1, 56, [128, 128], 64, True),
    ("deconv2", 1, 4, 2, 1, 112, [64, 64], 32, True),
    ("head_f", 0, 1, 1, 0, 224, [64], 32, False),
]
eng = ScnetEngine(None, mode='tc')
print('halo flags', eng.halo_flags)
eng._P, eng._dev, eng._bufs = G, dev, {'partials': None}
n = 2 * G
todo = []
for name, tr, k, s, p, Hin, cins, Cout, bn in LAYERS:
    srcs = []
    for c in cins:
        raw = torch.randn((n, Hin, Hin, c), device=dev).to(h16())
        srcs.append(_Act(raw, Hin, Hin, c, 0, c, torch.ones((G, c), device=dev), torch.zeros((G, c), device=dev)))
    Cin = sum(cins)
    Hout = (Hin - 1) * s - 2 * p + k if tr else (Hin + 2 * p - k) // s + 1
    eng._packed[name] = torch.randn((k, k, Cin, Cout), device=dev) / (Cin * k * k) ** 0.5
    if bn:
        out = _Act(torch.empty((n, Hout, Hout, Cout), device=dev, dtype=h16()), Hout, Hout, Cout, 0, Cout,
                   torch.zeros((G, Cout), device=dev), torch.zeros((G, Cout), device=dev))
        kw = dict(bn_params=(torch.ones(Cout, device=dev), torch.zeros(Cout, device=dev)))
    else:
        out = _Act(torch.empty((n, Hout, Hout, Cout), device=dev), Hout, Hout, Cout, 0, Cout)
        kw = dict(bn=False, bias=torch.zeros(Cout, device=dev))
    todo.append((name, srcs, out, bool(tr), k, s, p, kw))
st = torch.cuda.current_stream().cuda_stream
for it in range(1 + reps):
    if it == 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    for name, srcs, out, tr, k, s, p, kw in todo:
        if it == 0 or reps > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        eng._conv(name, srcs, out, tr, k, s, p, stream=st, **kw)
        if it == 0 or reps > 1:
            e1.record(); torch.cuda.synchronize()
            if it > 0 or reps == 1:
                print("%-8s %8.1f us (incl. bn_finalize)" % (name, e0.elapsed_time(e1) * 1e3))
                if eng.halo_flags & 32 and it == reps:
                    import ctypes
                    buf = (ctypes.c_ulonglong * 16)()
                    eng.lib.rp_conv_halo_prof(buf)
                    v = [float(x) for x in buf]
                    pc = lambda a, b: 100.0 * a / max(b, 1.0)
                    print("   loader: wait a_empty %.0f%%, copy issue %.0f%%, copy wait %.0f%%, transform %.0f%%, table %.0f%%, fence+arrive %.0f%% | mma: wait acc_empty %.0f%%, a_full %.0f%%, w_full %.0f%% | "
                          "epilogue: wait acc_full %.0f%%, tmem_ld %.0f%%, stats barriers+psum %.0f%%" % (
                              pc(v[1], v[0]), pc(v[13], v[0]), pc(v[2], v[0]), pc(v[3], v[0]), pc(v[12], v[0]), pc(v[14], v[0]), pc(v[5], v[4]), pc(v[6], v[4]), pc(v[7], v[4]),
                              pc(v[9], v[8]), pc(v[10], v[8]), pc(v[11], v[8])))
                elif eng.halo_flags & 32:
                    import ctypes
                    eng.lib.rp_conv_halo_prof((ctypes.c_ulonglong * 16)())
torch.cuda.synchronize(); torch.cuda.profiler.stop()

"""Generate tests/golden/rp_golden.npz by running the UNMODIFIED reference solver.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the reference ``RelativePoseEstimation_helper``
(RPModule/rpmodule.py:317-508, loaded through oracle/ref_loader.py) is executed
on seeded synthetic primitives (relativepose_b200/synth.py) and we record

  * ``T``        the returned 4x4 pose,
  * ``topk``     the index sets ``np.argpartition`` returned inside the helper
                 (rpmodule.py:369), sorted per row,
  * ``row/col/w`` the flat correspondence ids and pair weights handed to the
                 spectral fitters (rpmodule.py:495-505) -- i.e. the surviving
                 pair set after both filters -- when the method takes them,
  * the inputs themselves (so the vectors stay valid if the generator changes).

The reference holds no golden vectors of its own; these
files are what pins the oracle and the CUDA path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.ref_loader import load_reference_rpmodule, reference_opts  # noqa: E402
from relativepose_b200 import synth  # noqa: E402


class _NpSpy(object):
    """Stands in for the ``np`` global of the reference module and records argpartition."""

    def __init__(self):
        self.topk = None

    def __getattr__(self, name):
        return getattr(np, name)

    def argpartition(self, a, kth, axis=-1):
        r = np.argpartition(a, kth, axis=axis)
        self.topk = (r, kth)
        return r


def run_reference(ref, rec, para):
    spy = _NpSpy()
    grabbed = {}
    saved_np = ref.np
    saved_fits = {k: getattr(ref, k) for k in ('fit_irls_sm', 'fit_spectral')}

    def wrap(name):
        inner = saved_fits[name]

        def f(allSP, allTP, allSN, allTN, allWP, allWN, w, mu, row, col, ns, nt):
            grabbed['row'], grabbed['col'], grabbed['w'] = row.copy(), col.copy(), w.copy()
            return inner(allSP, allTP, allSN, allTN, allWP, allWN, w, mu, row, col, ns, nt)
        return f

    ref.np = spy
    for k in saved_fits:
        setattr(ref, k, wrap(k))
    try:
        s, t = synth.record_to_dicts(rec)
        T = ref.RelativePoseEstimation_helper(s, t, para)
    finally:
        ref.np = saved_np
        for k, v in saved_fits.items():
            setattr(ref, k, v)
    out = {'T': np.asarray(T, dtype=np.float64)}
    if spy.topk is not None:
        r, kth = spy.topk
        out['topk'] = np.sort(r[:, :kth], axis=1).astype(np.int32)
    for k in ('row', 'col', 'w'):
        if k in grabbed:
            out[k] = grabbed[k]
    return out


def cases():
    """(name, record, method, param_row, dataset, topK)"""
    out = []
    for n in (13, 26, 52, 103):
        for seed in (0, 1, 2):
            out.append(("n%d_s%d" % (n, seed), synth.make_pair(seed, n), 'irls+sm', 0, 'suncg', 5))
    out.append(("n205_s0", synth.make_pair(0, 205), 'irls+sm', 0, 'suncg', 5))
    # other parameter rows / datasets (evaluation.py:95-101 picks row alter_)
    out.append(("n52_s3_row1", synth.make_pair(3, 52), 'irls+sm', 1, 'suncg', 5))
    out.append(("n52_s4_row2", synth.make_pair(4, 52), 'irls+sm', 2, 'suncg', 5))
    out.append(("n52_s5_mp", synth.make_pair(5, 52), 'irls+sm', 0, 'matterport', 5))
    out.append(("n52_s6_scannet", synth.make_pair(6, 52), 'irls+sm', 0, 'scannet', 5))
    # the other three methods (rpmodule.py:491-500)
    for m in ('horn87', 'irls', 'spectral'):
        out.append(("n52_s7_%s" % m.replace('+', ''), synth.make_pair(7, 52), m, 0, 'suncg', 5))
        out.append(("n103_s8_%s" % m.replace('+', ''), synth.make_pair(8, 103), m, 0, 'suncg', 5))
    # ragged: n_s != n_t, low inlier ratio, noisy
    out.append(("rag_40_70", synth.make_pair(9, 40, 70), 'irls+sm', 0, 'suncg', 5))
    out.append(("rag_90_35", synth.make_pair(10, 90, 35), 'irls+sm', 0, 'suncg', 5))
    out.append(("low_inlier", synth.make_pair(11, 80, inlier_frac=0.2), 'irls+sm', 0, 'suncg', 5))
    out.append(("noisy", synth.make_pair(12, 80, pos_noise=0.03, feat_noise=0.1), 'irls+sm', 0, 'suncg', 5))
    # topK variants (rpmodule.py:368: topK=min(para.topK, n_t-1))
    out.append(("topk4_n64", synth.make_pair(13, 64), 'irls+sm', 0, 'suncg', 4))
    out.append(("topk_clamped", synth.make_pair(14, 30, 4), 'irls+sm', 0, 'suncg', 5))
    # descriptor arrays as transposed views, the layout the reference's own pipeline passes (rpmodule.py:531-532):
    # NumPy then sums the 32 squared differences sequentially instead of pairwise
    for nm, seed, n, which in (("forder_both", 18, 52, "st"), ("forder_src", 19, 40, "s"), ("forder_tgt_n103", 20, 103, "t")):
        r = synth.make_pair(seed, n)
        if "s" in which:
            r['feat_src'] = np.ascontiguousarray(r['feat_src'].T).T
        if "t" in which:
            r['feat_tgt'] = np.ascontiguousarray(r['feat_tgt'].T).T
        out.append((nm, r, 'irls+sm', 0, 'suncg', 5))
    # early exits (rpmodule.py:346-348, 406-408, 440-443)
    out.append(("exit_few_kp", synth.make_pair(15, 2, 10), 'irls+sm', 0, 'suncg', 5))
    out.append(("exit_no_inlier", synth.make_pair(16, 6, inlier_frac=0.0), 'irls+sm', 0, 'suncg', 5))
    r = synth.make_pair(17, 8)
    r['pc_tgt'] = r['pc_tgt'] * 1e-3       # all target points within the separation threshold
    out.append(("exit_collapsed_tgt", r, 'irls+sm', 0, 'suncg', 5))
    return out


def main():
    ref = load_reference_rpmodule()
    blob = {}
    names = []
    for name, rec, method, prow, ds, topk in cases():
        P = synth.shipped_params(ds)
        para = reference_opts(*P[prow])
        para.method = method
        para.topK = topk
        res = run_reference(ref, rec, para)
        names.append(name)
        blob[name + '/meta'] = np.array([P[prow][0], P[prow][1], P[prow][2], P[prow][3], topk], dtype=np.float64)
        blob[name + '/method'] = np.array(method)
        for k in ('pc_src', 'normal_src', 'feat_src', 'weight_src', 'pc_tgt', 'normal_tgt', 'feat_tgt',
                  'weight_tgt', 'R_gt'):
            blob[name + '/in/' + k] = rec[k]
        for k, v in res.items():
            blob[name + '/out/' + k] = v
        print("%-22s n_s=%4d n_t=%4d method=%-8s |T-Tgt|=%.2e pairs=%s" % (
            name, rec['pc_src'].shape[0], rec['pc_tgt'].shape[0], method,
            np.linalg.norm(res['T'] - rec['R_gt']), res['w'].shape[0] if 'w' in res else '-'))
    blob['names'] = np.array(names)
    path = os.path.join(HERE, 'rp_golden.npz')
    np.savez_compressed(path, **blob)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == '__main__':
    main()

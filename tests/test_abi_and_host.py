"""CPU-side checks: the C-ABI library loads and exports every symbol include/rp_b200.h declares; host logic."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "rp_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(rp_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from relativepose_b200 import _lib
    lib = _lib.load()
    names = _header_functions()
    assert "rp_solve_batch" in names and "rp_match_topk" in names
    for n in names:
        assert hasattr(lib, n), "librp_b200.so does not export %s" % n
    assert lib.rp_abi_version() == 1
    assert set(_lib.EXPORTS) <= set(names)


def test_struct_layouts_match_header():
    from relativepose_b200 import _lib
    assert ctypes.sizeof(_lib.RpParams) == 10 * 8 + 4 * 4
    assert _lib.RpParams.topk.offset == 80 and _lib.RpParams.method.offset == 84
    assert ctypes.sizeof(_lib.RpDebug) == 10 * 8


def test_params_follow_reference_expressions():
    from relativepose_b200.RPModule.rputil import opts
    from relativepose_b200.solver import params_from_opts
    para = opts(0.3, 0.4, 0.05, 0.009)
    p = params_from_opts(para)
    sig = np.ones([1]) * 0.009
    assert p.feat_den == float((2 * np.power(sig / 5, 2))[0])
    assert p.feat_den_obs == float((2 * np.power((np.ones([1]) * (0.009 / 1.2)) / 5, 2))[0])
    assert p.dist_thre_sq == float(np.power(0.08, 2)) and p.sep_thre == float(1.5 * np.power(1.5 * 0.08, 2))
    assert p.angle_thre_sq == float(np.power(45 / 180. * np.pi, 2))
    assert p.den_dist == 2 * 0.05 ** 2 and p.den_a1 == 2 * 0.3 ** 2 and p.den_a2 == 2 * 0.4 ** 2
    assert p.mu == 0.3 and p.topk == 5 and p.method == 3
    para.method = "what"
    with pytest.raises(Exception):
        params_from_opts(para)


def test_opts_defaults_match_reference_fields():
    from RPModule.rputil import opts
    o = opts()
    assert (o.distThre, o.distSepThre, o.mu, o.topK, o.method) == (0.08, 1.5 * 0.08, 0.3, 5, 'irls+sm')
    assert o.sigmaAngle1 == 0.523 / 2 and o.sigmaDist == 0.04 and o.sigmaFeat == 0.01
    assert abs(o.angleThre - np.pi / 4) < 1e-15


def test_zero_row_table_is_numpy_tie_order():
    from relativepose_b200.solver import PackedBatch, zero_row_topk
    from relativepose_b200 import synth
    for n_t, K in ((4, 3), (8, 5), (103, 5), (77, 4)):
        ref = np.argpartition(-np.zeros([3, n_t]), K, axis=1)[:, :K]
        assert np.array_equal(zero_row_topk(n_t, K), ref[0]) and np.array_equal(ref[0], ref[2])
    pk = PackedBatch([synth.make_pair(1, 6, 9), synth.make_pair(2, 5, 4)], pin=False)
    tab = pk.zero_rows(5, 5).numpy()
    assert tab.shape == (2, 5) and (tab[1, 3:] == -1).all() and (tab[0] >= 0).all()


def test_packed_batch_layout():
    from relativepose_b200.solver import PackedBatch
    from relativepose_b200 import synth
    recs = [synth.make_pair(3, 7, 11), synth.make_pair(4, 13, 5)]
    pk = PackedBatch(recs, pin=False)
    assert pk.B == 2 and list(pk.off_s) == [0, 7, 20] and list(pk.off_t) == [0, 11, 16]
    assert pk.max_ns == 13 and pk.max_nt == 11 and pk.feat_dim == 32
    assert pk.feat_s.dtype.is_floating_point and pk.feat_s.element_size() == 4 and pk.pc_s.element_size() == 8
    assert np.array_equal(pk.pc_t.numpy()[11:], recs[1]["pc_tgt"])


def test_product_fails_loudly_without_cuda():
    """No CPU fallback: constructing the solver without a GPU raises."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from relativepose_b200.solver import PoseSolver
    with pytest.raises(RuntimeError):
        PoseSolver()


def test_product_never_imports_oracle():
    bad = []
    for dp, _, fs in os.walk(os.path.join(ROOT, "relativepose_b200")):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M):
                    bad.append(f)
    assert not bad, "product modules import the oracle: %s" % bad


def test_round2_entry_points_validate_their_arguments_without_a_gpu():
    """rp_solve_pair_host rejects null / out-of-range arguments before it touches the device; rp_solver_wide_max is a plain
    get / set (default: the SM count, 148 when no device can be asked)."""
    import ctypes
    from relativepose_b200 import _lib
    lib = _lib.load()
    p = _lib.RpParams()
    T = (ctypes.c_double * 16)()
    st = (ctypes.c_int32 * 1)()
    z = ctypes.c_void_p(0)
    buf = (ctypes.c_double * 64)()
    b = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.rp_solve_pair_host(10, 10, z, z, z, z, z, z, z, z, 32, ctypes.byref(p), z, 5, 0, 0, ctypes.cast(T, ctypes.c_void_p),
                                ctypes.cast(st, ctypes.c_void_p), z, z)
    assert rc == -1          # RP_ERR_INVALID_ARG
    rc = lib.rp_solve_pair_host(10, 10, b, b, b, b, b, b, b, b, 32, ctypes.byref(p), z, _lib.MAX_TOPK + 1, 0, 0,
                                ctypes.cast(T, ctypes.c_void_p), ctypes.cast(st, ctypes.c_void_p), z, z)
    assert rc == -3          # RP_ERR_UNSUPPORTED
    old = lib.rp_solver_wide_max(-1)
    assert old >= 0
    assert lib.rp_solver_wide_max(7) == old and lib.rp_solver_wide_max(-1) == 7
    lib.rp_solver_wide_max(old)
    assert lib.rp_solver_wide_max(-1) == old

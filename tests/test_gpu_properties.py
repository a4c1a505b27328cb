"""Size-independent properties of the CUDA solver (test-pyramid items 2 and 4) and the stage entries
that had only symbol-presence coverage.

  * rigid-motion equivariance: moving the target scan by G moves the pose to G T; moving the source by G gives T G^-1
    (distances and angles are invariant, Horn's fit is equivariant; only float rounding differs);
  * keypoint permutation: shuffling the rows of either scan leaves the pose unchanged;
  * sharding: the same pair list solved whole, in 2 / 4 / 8 contiguous shards (relativepose_b200.sharding.shard_bounds) and
    with a different number of resident CTAs gives BITWISE identical poses per pair -- on one GPU, and across two GPUs
    when the box has them;
  * rp_match_topk (stage entry, rpmodule.py:342-375): index sets and soft-match weights against the oracle.
"""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _para():
    from relativepose_b200 import synth
    from relativepose_b200.RPModule.rputil import opts
    return opts(*synth.shipped_params("suncg")[0])


def _rigid(seed):
    from relativepose_b200 import synth
    return synth.make_pose(seed, max_angle=2.5, t_sigma=1.0)


def _move(rec, side, G):
    r = dict(rec)
    R, t = G[:3, :3], G[:3, 3]
    r['pc_' + side] = rec['pc_' + side] @ R.T + t
    r['normal_' + side] = rec['normal_' + side] @ R.T
    return r


@pytest.mark.parametrize("seed", [3, 14, 15])
def test_rigid_motion_equivariance(seed):
    from relativepose_b200 import synth
    from RPModule.rpmodule import RelativePoseEstimation_batch
    rec = synth.make_pair(seed, 80, 90)
    G = _rigid(seed)
    T, Tt, Ts = RelativePoseEstimation_batch([rec, _move(rec, 'tgt', G), _move(rec, 'src', G)], _para())
    assert not np.array_equal(T, np.eye(4))
    assert np.linalg.norm(Tt - G @ T) <= 1e-9 * max(1.0, np.linalg.norm(T))
    assert np.linalg.norm(Ts - T @ np.linalg.inv(G)) <= 1e-9 * max(1.0, np.linalg.norm(T))


@pytest.mark.parametrize("seed", [2, 7])
def test_keypoint_permutation_invariance(seed):
    from relativepose_b200 import synth
    from RPModule.rpmodule import RelativePoseEstimation_batch
    rec = synth.make_pair(seed, 70, 64)
    rs = np.random.RandomState(seed)
    ps, pt = rs.permutation(70), rs.permutation(64)
    r2 = dict(rec)
    for k in ('pc', 'normal', 'feat', 'weight'):
        r2[k + '_src'] = np.ascontiguousarray(rec[k + '_src'][ps])
        r2[k + '_tgt'] = np.ascontiguousarray(rec[k + '_tgt'][pt])
    T, T2 = RelativePoseEstimation_batch([rec, r2], _para())
    assert not np.array_equal(T, np.eye(4))
    assert np.linalg.norm(T - T2) <= 1e-10


def test_sharded_solves_are_bitwise_identical():
    import torch
    from relativepose_b200 import sharding, synth
    from relativepose_b200.solver import PoseSolver
    rs = np.random.RandomState(0)
    recs = [synth.make_pair(500 + i, int(rs.randint(20, 90)), int(rs.randint(20, 90))) for i in range(37)]
    para = _para()
    whole = PoseSolver("cuda:0").solve_records(recs, para)
    for world in (2, 4, 8):
        parts = []
        for rank in range(world):
            lo, hi = sharding.shard_bounds(len(recs), rank, world)
            dev = "cuda:%d" % (rank % torch.cuda.device_count())          # real second GPU when the box has one
            parts.append(PoseSolver(dev, n_slots=1 + rank).solve_records(recs[lo:hi], para))
        assert np.array_equal(np.concatenate(parts, 0), whole), "world=%d" % world
    one_by_one = np.stack([PoseSolver("cuda:0").solve_records([r], para)[0] for r in recs[:8]])
    assert np.array_equal(one_by_one, whole[:8])


def test_match_topk_stage_entry():
    """rp_match_topk through the C ABI: per-row index sets identical to the oracle's (= numpy.argpartition on the
    reference's wij), soft-match weights of the selected entries to 1e-12."""
    import torch
    from oracle import rp_oracle
    from relativepose_b200 import _lib, synth
    from relativepose_b200.solver import PackedBatch, default_solver, params_from_opts
    lib = _lib.load()
    recs = [synth.make_pair(900 + i, n_s, n_t) for i, (n_s, n_t) in enumerate(((40, 36), (25, 60), (33, 33)))]
    para = _para()
    sol = default_solver('cuda:0')
    pk = PackedBatch(recs)
    d = pk.to_device(sol.device)
    K = 5
    ws, key = sol._workspace(pk.B, d.max_ns, d.max_nt, K, d.feat_dim, 0)
    par = sol._params_device([params_from_opts(para)])
    tot = int(pk.off_s[-1])
    idx = torch.full((tot, K), -9, dtype=torch.int32, device=sol.device)
    f = torch.zeros((tot, K), dtype=torch.float64, device=sol.device)
    status = torch.full((pk.B,), -9, dtype=torch.int32, device=sol.device)
    zr = d.zero_rows(K, K, sol.device)
    rc = lib.rp_match_topk(pk.B, d.off_s_t.data_ptr(), d.off_t_t.data_ptr(), d.feat_s.data_ptr(), d.w_s.data_ptr(),
                           d.feat_t.data_ptr(), d.w_t.data_ptr(), d.feat_dim, par.data_ptr(), None, zr.data_ptr(),
                           d.sum_order_t.data_ptr(), key[0], key[1], K, key[5], ws.data_ptr(), ws.numel(),
                           idx.data_ptr(), f.data_ptr(), status.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    assert (status.cpu().numpy() == 0).all()
    idx, f = idx.cpu().numpy(), f.cpu().numpy()
    from relativepose_b200 import synth as _s
    for b, rec in enumerate(recs):
        s, t = _s.record_to_dicts(rec)
        tr = {}
        rp_oracle.solve_pair(s, t, rp_oracle.Params(*_s.shipped_params("suncg")[0]), tr)
        lo, hi = int(pk.off_s[b]), int(pk.off_s[b + 1])
        got = idx[lo:hi]
        assert np.array_equal(np.sort(got, axis=1), tr['topk'])
        wij = tr['wij']
        ref_f = np.take_along_axis(wij, got.astype(np.int64), axis=1)
        assert np.allclose(f[lo:hi], ref_f, rtol=1e-12, atol=1e-300)


def test_ambiguous_topk_rows_are_counted():
    """stats[6] >> 8 counts the source rows whose top-k set is not decided by the keys (include/rp_b200.h): zero on the
    generic synthetic pairs, non-zero when target descriptors are exact duplicates that straddle the top-k boundary."""
    from relativepose_b200 import synth
    from relativepose_b200.solver import PoseSolver
    para = _para()
    sol = PoseSolver("cuda:0")
    recs = [synth.make_pair(40 + i, 50, 60) for i in range(4)]
    _, st, stats = sol.solve_records(recs, para, return_stats=True)
    assert (st == 0).all() and ((stats[:, 6] >> 8) == 0).all()
    r = dict(recs[0])
    ft = r['feat_tgt'].copy()
    ft[:] = r['feat_src'][0]                       # every target descriptor equals source row 0: all its keys tie at 0
    r['feat_tgt'] = ft
    _, st2, stats2 = sol.solve_records([r, recs[1]], para, return_stats=True)
    assert (stats2[0, 6] >> 8) > 0 and (stats2[1, 6] >> 8) == 0

"""GPU parity tests for the fused solver: CUDA path (through the C ABI) vs golden vectors from the
reference and vs the numpy oracle on seeded inputs.

Bars (BASELINE.json north_star): top-k index sets bit-exact, surviving-pair set identical,
||T - T_ref||_F <= 1e-4 (we assert a much tighter 1e-8 on these well-conditioned cases)."""
import ctypes

import numpy as np
import pytest

from tests.golden_util import load_cases

pytestmark = pytest.mark.gpu

CASES = load_cases()
# near-bipartite affinity (+lambda / -lambda of equal magnitude): scipy's ARPACK starts from a random vector, so the
# reference itself returns either eigenvector from run to run; only the discrete stages are compared for it
ILL_POSED = ('topk_clamped',)
T_TOL = 1e-8


@pytest.fixture(scope="module")
def solver():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from relativepose_b200.solver import PoseSolver
    return PoseSolver("cuda:0")


def _debug_run(solver, records, para, edge_cap=20000, stop_after=3):
    import torch
    from relativepose_b200 import _lib
    from relativepose_b200.solver import PackedBatch, params_from_opts
    pk = PackedBatch(records)
    d = pk.to_device(solver.device)
    B = pk.B
    K = max(1, min(int(para.topK), pk.max_nt - 1))   # = the max_topk stride solve_device passes
    tot_s = int(pk.off_s[-1])
    dev = solver.device
    topk_idx = torch.full((tot_s, K), -7, dtype=torch.int32, device=dev)
    topk_f = torch.zeros((tot_s, K), dtype=torch.float64, device=dev)
    sizes = [(int(pk.off_s[b + 1] - pk.off_s[b])) * int(pk.off_t[b + 1] - pk.off_t[b]) for b in range(B)]
    dij_off = torch.tensor(np.concatenate(([0], np.cumsum(sizes)[:-1])).astype(np.int64), device=dev)
    dij = torch.zeros((int(sum(sizes)),), dtype=torch.float32, device=dev)
    edge_rc = torch.full((B, edge_cap, 2), -1, dtype=torch.int32, device=dev)
    edge_w = torch.zeros((B, edge_cap), dtype=torch.float64, device=dev)
    dbg = _lib.RpDebug()
    dbg.topk_idx = topk_idx.data_ptr(); dbg.topk_f = topk_f.data_ptr()
    dbg.dij = dij.data_ptr(); dbg.dij_off = dij_off.data_ptr()
    dbg.edge_rc = edge_rc.data_ptr(); dbg.edge_w = edge_w.data_ptr(); dbg.edge_cap = edge_cap
    dbg.u = None; dbg.u_stride = 0
    plist = [params_from_opts(para)]
    # topk passed to the library must be the *parameter's* topK (max over batch n_t handled inside)
    T, status, stats = solver.solve_device(d, plist, stop_after=stop_after, debug=dbg)
    torch.cuda.synchronize()
    return dict(T=T.cpu().numpy(), status=status.cpu().numpy(), stats=stats.cpu().numpy(),
                topk_idx=topk_idx.cpu().numpy(), topk_f=topk_f.cpu().numpy(), dij=dij.cpu().numpy(),
                dij_off=dij_off.cpu().numpy(), edge_rc=edge_rc.cpu().numpy(), edge_w=edge_w.cpu().numpy(), pk=pk)


@pytest.mark.parametrize("case", CASES, ids=[c.name for c in CASES])
def test_cuda_matches_reference_golden(solver, case):
    from relativepose_b200.RPModule.rputil import opts
    from oracle import rp_oracle
    para = case.apply(opts())
    out = _debug_run(solver, [case.record], para)
    s, t = case.dicts()
    n_s, n_t = s['pc'].shape[0], t['pc'].shape[0]
    trace = {}
    T_or = rp_oracle.solve_pair(s, t, case.apply(rp_oracle.Params()), trace)
    assert int(out['status'][0]) == trace['status']
    if n_s >= 3 and n_t >= 3:
        K = int(out['stats'][0, 7]) & 0xff
        # float32 descriptor distances: bit exact vs numpy
        assert np.array_equal(out['dij'][:n_s * n_t].reshape(n_s, n_t), trace['dij'])
        got = np.sort(out['topk_idx'][:n_s, :K], axis=1)
        assert np.array_equal(got, case.topk_sets), "top-k index sets differ from the reference"
    if case.row is not None:
        M = int(out['stats'][0, 2])
        assert M == case.row.shape[0]
        K = int(out['stats'][0, 7]) & 0xff
        cj = out['topk_idx'][:n_s, :K].reshape(-1)
        ci = np.repeat(np.arange(n_s), K)
        flat = ci * n_t + cj
        rc = out['edge_rc'][0, :M]
        got = np.stack((flat[rc[:, 0]], flat[rc[:, 1]]), 1)
        # canonical order for set comparison
        def canon(a):
            a = np.sort(a, axis=1)
            return a[np.lexsort((a[:, 1], a[:, 0]))]
        ref = np.stack((case.row, case.col), 1)
        assert np.array_equal(canon(got), canon(ref)), "surviving pair set differs from the reference"
        # weights: match pair by pair
        order_g = np.lexsort((np.sort(got, 1)[:, 1], np.sort(got, 1)[:, 0]))
        order_r = np.lexsort((np.sort(ref, 1)[:, 1], np.sort(ref, 1)[:, 0]))
        assert np.allclose(out['edge_w'][0, :M][order_g], case.w[order_r], rtol=1e-11, atol=0)
    err = np.linalg.norm(out['T'][0] - case.T)
    if case.name not in ILL_POSED:
        assert err <= T_TOL, "||T - T_ref||_F = %g" % err
    # Live oracle on this host: only meaningful where the oracle reproduces the reference's golden pose here
    # (ARPACK may return the -lambda eigenvector of a near-bipartite affinity on another CPU: 'topk_clamped').
    if case.name in ILL_POSED:
        # bipartite affinity: lambda_max = |lambda_min|, the reference's eigs(k=1) = 'largest magnitude' returns either end
        # of the spectrum depending on the platform (the golden has the -lambda vector in alternations 1, 3, 5).  The CUDA
        # path always returns the Perron vector: compare with the oracle asked for the largest REAL part.
        rp_oracle.EIG_WHICH = 'LR'
        try:
            T_lr = rp_oracle.solve_pair(s, t, case.apply(rp_oracle.Params()))
        finally:
            rp_oracle.EIG_WHICH = 'LM'
        print("%s: |T - T_oracle(LR)| = %.2e, |T - golden| = %.2e, eigen its %s" % (case.name, np.linalg.norm(out['T'][0] - T_lr), err, out['stats'][0, 4:7]))
        assert np.linalg.norm(out['T'][0] - T_lr) <= T_TOL
    elif np.linalg.norm(T_or - case.T) <= T_TOL:
        assert np.linalg.norm(out['T'][0] - T_or) <= T_TOL


def test_batch_ragged_matches_oracle(solver):
    from relativepose_b200 import synth
    from relativepose_b200.RPModule.rputil import opts
    from oracle import rp_oracle
    P = synth.shipped_params('suncg')
    recs = [synth.make_pair(200 + i, n_s, n_t) for i, (n_s, n_t) in
            enumerate([(20, 31), (64, 64), (2, 9), (45, 17), (103, 103), (8, 8), (33, 90)])]
    para = opts(*P[0])
    T, status, stats = solver.solve_records(recs, para, return_stats=True)
    for b, r in enumerate(recs):
        s, t = synth.record_to_dicts(r)
        tr = {}
        To = rp_oracle.solve_pair(s, t, rp_oracle.Params(*P[0]), tr)
        assert status[b] == tr['status']
        assert np.linalg.norm(T[b] - To) <= T_TOL, (b, np.linalg.norm(T[b] - To))


def test_helper_api_single_pair(solver):
    from RPModule.rpmodule import RelativePoseEstimation_helper
    from RPModule.rputil import opts
    c = [c for c in CASES if c.name == 'n52_s0'][0]
    s, t = c.dicts()
    T = RelativePoseEstimation_helper(s, t, c.apply(opts()))
    assert T.shape == (4, 4) and T.dtype == np.float64
    assert np.linalg.norm(T - c.T) <= T_TOL
    para = c.apply(opts()); para.method = 'bogus'
    with pytest.raises(Exception):
        RelativePoseEstimation_helper(s, t, para)
    assert np.array_equal(RelativePoseEstimation_helper({k: v[:2] for k, v in s.items()}, t, c.apply(opts())), np.eye(4))


def test_determinism_and_slot_independence(solver):
    """Same pairs through different slot counts / batch positions give bitwise identical poses."""
    from relativepose_b200 import synth
    from relativepose_b200.solver import PoseSolver
    from relativepose_b200.RPModule.rputil import opts
    P = synth.shipped_params('suncg')
    recs = synth.make_batch(300, 24, 52)
    para = opts(*P[0])
    T1 = solver.solve_records(recs, para)
    T2 = PoseSolver("cuda:0", n_slots=3).solve_records(recs[::-1], para)[::-1]
    T3 = solver.solve_records(recs, para)
    assert np.array_equal(T1, T3)
    assert np.array_equal(T1, T2)


def test_pipelined_chunks_bitwise_equal(solver):
    """solve_packed cuts large batches into chunks to overlap H2D with compute: same bits as one launch."""
    from relativepose_b200 import synth
    from relativepose_b200.solver import PackedBatch
    from relativepose_b200.RPModule.rputil import opts
    P = synth.shipped_params('suncg')
    recs = [synth.make_pair(700 + i, n_s, n_t) for i, (n_s, n_t) in enumerate([(30, 41), (52, 52), (2, 9), (64, 33), (26, 26)] * 4)]
    pk = PackedBatch(recs)
    para = opts(*P[0])
    T1, s1, st1 = solver.solve_packed(pk, para, return_stats=True, chunks=1)
    T4, s4, st4 = solver.solve_packed(pk, para, return_stats=True, chunks=4)
    assert np.array_equal(T1, T4) and np.array_equal(s1, s4) and np.array_equal(st1, st4)


def test_large_n_sweep_end(solver):
    """configs[4] upper end: nominal N=2048 (n_s=n_t=410, N_actual=2050) against the oracle."""
    from relativepose_b200 import synth
    from relativepose_b200.RPModule.rputil import opts
    from oracle import rp_oracle
    P = synth.shipped_params('suncg')
    rec = synth.make_pair(4242, 410)
    T, status, stats = solver.solve_records([rec], opts(*P[0]), return_stats=True)
    s, t = synth.record_to_dicts(rec)
    tr = {}
    To = rp_oracle.solve_pair(s, t, rp_oracle.Params(*P[0]), tr)
    assert status[0] == 0 and stats[0, 0] == 2050
    assert stats[0, 2] == tr['pairs'].shape[0], "surviving pair count differs"
    assert np.linalg.norm(T[0] - To) <= T_TOL


def test_small_batch_path_is_bitwise_the_general_path(solver):
    """PoseSolver.solve_records sends <= 16 pairs through one staging buffer / one H2D / one D2H (_solve_small); the poses,
    status and stats must be bit-identical to the general pinned-array path, for ragged pairs and both descriptor layouts."""
    from relativepose_b200 import synth
    from relativepose_b200.RPModule.rputil import opts
    from relativepose_b200.solver import PackedBatch
    recs = [synth.make_pair(300 + i, n, m) for i, (n, m) in enumerate(((26, 31), (40, 40), (52, 37), (3, 9), (2, 5)))]
    recs[1] = dict(recs[1]); recs[1]['feat_src'] = np.asfortranarray(recs[1]['feat_src'])      # a transposed-view layout
    para = opts(*synth.shipped_params('suncg')[0])
    Ts, ss, sts = solver.solve_records(recs, para, return_stats=True)
    Tg, sg, stg = solver.solve_packed(PackedBatch(recs), para, return_stats=True)
    assert np.array_equal(Ts, Tg) and np.array_equal(ss, sg) and np.array_equal(sts, stg)
    T1 = solver.solve_records(recs[:1], para)
    assert np.array_equal(T1[0], Tg[0])


def test_wide_build_matches_the_default_build(solver):
    """The 512-thread build (small batches, rp_solver_wide_max) against the 128-thread build and the oracle: identical status,
    pair counts and top-k decisions (stats), poses equal to rounding (reduction trees differ with the CTA width), and the wide
    build is itself deterministic (bitwise) across slot counts and batch positions."""
    from oracle import rp_oracle
    from relativepose_b200 import _lib, synth
    from relativepose_b200.RPModule.rputil import opts
    from relativepose_b200.solver import PoseSolver
    lib = _lib.load()
    P = synth.shipped_params('suncg')
    para = opts(*P[0])
    recs = [synth.make_pair(900 + i, n, m) for i, (n, m) in enumerate(((103, 103), (60, 140), (26, 31), (3, 9), (2, 5), (88, 45)))]
    old = lib.rp_solver_wide_max(-1)
    try:
        lib.rp_solver_wide_max(0)
        Tn, sn, stn = solver.solve_records(recs, para, return_stats=True)
        lib.rp_solver_wide_max(1 << 20)
        Tw, sw, stw = solver.solve_records(recs, para, return_stats=True)
        Tw2 = PoseSolver("cuda:0", n_slots=2).solve_records(recs[::-1], para)[::-1]
    finally:
        lib.rp_solver_wide_max(old)
    assert np.array_equal(sn, sw)
    assert np.array_equal(stn[:, :4], stw[:, :4]) and np.array_equal(stn[:, 6:], stw[:, 6:])      # N, M1, M2, NZ, flags, K
    assert np.abs(Tn - Tw).max() <= 1e-11
    assert np.array_equal(Tw, Tw2)
    for rec, T in zip(recs, Tw):
        s, t = synth.record_to_dicts(rec)
        assert np.linalg.norm(T - rp_oracle.solve_pair(s, t, rp_oracle.Params(*P[0]))) <= T_TOL

"""Host-side pieces that need no GPU: the primitive-cache format of the FD trainer, dataset ids, the op-list ABI struct."""
import ctypes
import os

import numpy as np
import pytest


def test_primitive_cache_roundtrip(tmp_path):
    from relativepose_b200 import fd_objective, synth
    prims = [synth.make_pair(5 + i, 20 + i) for i in range(3)]
    path = os.path.join(str(tmp_path), "primitives.npy")
    fd_objective.save_primitives(path, prims)                  # trainRelativePoseModuleRecFD.py:212
    back = fd_objective.load_primitives(path)                  # :113
    assert len(back) == 3
    for a, b in zip(prims, back):
        assert set(a) == set(b)
        for k in a:
            assert np.array_equal(a[k], b[k]) and np.asarray(a[k]).dtype == np.asarray(b[k]).dtype


def test_dataset_ids_follow_the_reference_substring_rule():
    from relativepose_b200 import util
    assert util.dataset_id('suncg') == 0 and util.dataset_id('matterport') == 1 and util.dataset_id('scannet') == 2
    assert util.dataset_id('data/suncg/test') == 0            # the reference tests `'suncg' in dataList` (util.py:97)
    with pytest.raises(ValueError):
        util.dataset_id('nyu')


def test_net_op_struct_matches_the_header_layout():
    """rp_net_op (include/rp_b200.h): int32 kind, int32 reserved, rp_conv_desc, 16 x uint64."""
    from relativepose_b200 import _lib
    assert ctypes.sizeof(_lib.RpNetOp) == 8 + ctypes.sizeof(_lib.RpConvDesc) + 16 * 8
    assert _lib.RpNetOp.conv.offset == 8 and _lib.RpNetOp.arg.offset == 8 + ctypes.sizeof(_lib.RpConvDesc)
    assert ctypes.sizeof(_lib.RpConvDesc) % 8 == 0
    assert sorted(_lib.NET_OPS.values()) == [1] + list(range(3, 16))      # 2 was the per-tap tensor-core kernel (removed in round 2)


def test_small_batch_constants():
    from relativepose_b200.solver import PoseSolver
    assert 1 <= PoseSolver.SMALL_BATCH <= 64

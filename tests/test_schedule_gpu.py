"""Invariants of the ticket schedule the TMA-staged AdaIN kernel walks (prologue / steady state / epilogue with
twin items, merge lead and lag), dumped through the kernel's REAL decode function for call shapes no data test
covers: few planes (< lag), odd chunk counts, 8 MiB planes, every style/prev combination.  A schedule that
violated these would hang (an apply waiting for a merge that is ticketed later) or drop data."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dump(planes, hw, has_style, has_prev, stats_only):
    import rpst
    D = rpst._lib.debug_lib()     # white-box hooks live in librpst_debug.so, not in the product library
    info = (ctypes.c_int64 * 5)()
    rpst._lib.check(D.rpst_debug_adain_schedule(planes, hw, has_style, has_prev, stats_only, None, 0, info, None))
    total = int(info[0])
    buf = torch.empty(total, 3, dtype=torch.int32, device="cuda")
    rpst._lib.check(D.rpst_debug_adain_schedule(planes, hw, has_style, has_prev, stats_only, buf.data_ptr(), total, info,
                                                torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return buf.cpu().numpy(), [int(v) for v in info]


@pytest.mark.parametrize("planes", [1, 2, 3, 4, 7, 31, 32, 33, 100])
@pytest.mark.parametrize("hw", [16388, 65536, 90000, 262144, 1100000, 2097152])
@pytest.mark.parametrize("has_style,has_prev,stats_only", [(1, 0, 0), (1, 1, 0), (0, 0, 0), (0, 0, 1), (1, 0, 1)])
def test_schedule_invariants(planes, hw, has_style, has_prev, stats_only):
    t, (total, ips, ipa, lag, lead) = dump(planes, hw, has_style, has_prev, stats_only)
    kind, plane, chunk = t[:, 0], t[:, 1], t[:, 2]
    idx = np.arange(total)
    assert set(np.unique(kind)) <= {0, 1, 2} and plane.min() >= 0 and plane.max() == planes - 1
    ipp = -(-hw // 4096)
    assert ips == (ipp if has_style else (ipp + 1) // 2)
    if stats_only:
        assert total == planes * ips and (kind == 0).all()
        assert np.array_equal(plane, idx // ips) and np.array_equal(chunk, idx % ips)
        return
    assert ipa == (ipp if has_prev else (ipp + 1) // 2) and 1 <= lead <= max(lag - 1, 1)
    assert total == planes * (ips + ipa + 1)
    for k, per in ((0, ips), (1, ipa), (2, 1)):
        sel = kind == k
        # every (plane, chunk) exactly once, chunks 0..per-1 in order inside a plane, planes in order
        assert sel.sum() == planes * per
        assert np.array_equal(plane[sel], np.repeat(np.arange(planes), per))
        assert np.array_equal(chunk[sel], np.tile(np.arange(per), planes) if k != 2 else np.zeros(planes, dtype=np.int32))
    last_stat = np.array([idx[(kind == 0) & (plane == p)].max() for p in range(planes)])
    merge_at = np.array([idx[(kind == 2) & (plane == p)][0] for p in range(planes)])
    first_apply = np.array([idx[(kind == 1) & (plane == p)].min() for p in range(planes)])
    assert (last_stat < merge_at).all() and (merge_at < first_apply).all()     # no apply can wait for a later ticket
    # the statistics of plane p + lag - 1 (or the last plane) are ticketed before plane p is applied: the lag that
    # keeps the content of `lag` planes between its two reads
    ahead = np.minimum(np.arange(planes) + lag - 1, planes - 1)
    assert (last_stat[ahead] < first_apply).all()


def dump_seg(n, c, hw_c, hw_s, has_prev):
    import rpst
    D = rpst._lib.debug_lib()     # white-box hooks live in librpst_debug.so, not in the product library
    info = (ctypes.c_int64 * 5)()
    rpst._lib.check(D.rpst_debug_seg_schedule(n, c, hw_c, hw_s, has_prev, None, 0, info, None))
    total = int(info[0])
    buf = torch.empty(total, 3, dtype=torch.int32, device="cuda")
    rpst._lib.check(D.rpst_debug_seg_schedule(n, c, hw_c, hw_s, has_prev, buf.data_ptr(), total, info,
                                              torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return buf.cpu().numpy(), [int(v) for v in info]


@pytest.mark.parametrize("n,c", [(1, 1), (1, 2), (2, 3), (1, 40), (3, 11)])
@pytest.mark.parametrize("hw_c,hw_s", [(4096, 4096), (65536, 40000), (90000, 262144), (2097152, 2097152)])
@pytest.mark.parametrize("has_prev", [0, 1])
def test_seg_schedule_invariants(n, c, hw_c, hw_s, has_prev):
    """Segment kernel: content statistics (0), style statistics (1), merge (3), apply (2) of every plane exactly
    once, statistics < merge < apply per plane, apply trailing by the lag."""
    t, (total, ic, is_, ia, lag) = dump_seg(n, c, hw_c, hw_s, has_prev)
    planes = n * c
    kind, plane, chunk = t[:, 0], t[:, 1], t[:, 2]
    idx = np.arange(total)
    assert ic == -(-hw_c // 4096) and is_ == -(-hw_s // 4096) and ia == -(-hw_c // (2048 if has_prev else 4096))
    assert total == planes * (ic + is_ + 1 + ia) and set(np.unique(kind)) == {0, 1, 2, 3}
    for k, per in ((0, ic), (1, is_), (2, ia), (3, 1)):
        sel = kind == k
        assert sel.sum() == planes * per
        assert np.array_equal(plane[sel], np.repeat(np.arange(planes), per))
        assert np.array_equal(chunk[sel], np.tile(np.arange(per), planes) if k != 3 else np.zeros(planes, dtype=np.int32))
    last_stat = np.array([idx[(kind <= 1) & (plane == p)].max() for p in range(planes)])
    merge_at = np.array([idx[(kind == 3) & (plane == p)][0] for p in range(planes)])
    first_apply = np.array([idx[(kind == 2) & (plane == p)].min() for p in range(planes)])
    assert (last_stat < merge_at).all() and (merge_at < first_apply).all()
    ahead = np.minimum(np.arange(planes) + lag - 1, planes - 1)
    assert (last_stat[ahead] < first_apply).all()

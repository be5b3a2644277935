"""Delay-schedule known-answer vectors measured from the reference object code
(Corr::calculateLevelMax / Corr::delaysPerLevel, corr.cpp:1133-1160)."""
import numpy as np
import pytest

KATS = [  # F, dpl, maxLevel, T, last (level, tau), per-level counts
    (15, 8, 0, 14, (0, 14), [14]),
    (16, 8, 0, 15, (0, 15), [15]),
    (17, 8, 0, 16, (0, 16), [16]),
    (33, 8, 1, 23, (1, 30), [16, 7]),
    (600, 8, 6, 56, (5, 512), [16, 8, 8, 8, 8, 8, 0]),
    (1000, 8, 6, 62, (6, 896), [16, 8, 8, 8, 8, 8, 6]),
    (9999, 8, 10, 88, (9, 8192), [16] + [8] * 9 + [0]),
    (10000, 8, 10, 88, (9, 8192), [16] + [8] * 9 + [0]),
    (10000, 4, 10, 48, (10, 8192), [8] + [4] * 10),
    (20000, 8, 11, 96, (10, 16384), [16] + [8] * 10 + [0]),
    (100000, 8, 13, 115, (13, 90112), [16] + [8] * 12 + [3]),
    (1000000, 8, 16, 142, (16, 917504), [16] + [8] * 15 + [6]),
]
F600 = (list(range(1, 17)) + list(range(18, 33, 2)) + list(range(36, 65, 4)) + list(range(72, 129, 8))
        + list(range(144, 257, 16)) + list(range(288, 513, 32)))


@pytest.mark.parametrize("F,dpl,ml,T,last,counts", KATS)
def test_oracle_schedule_kats(oracle, F, dpl, ml, T, last, counts):
    assert oracle.level_max(F, dpl) == ml
    lv, tv = oracle.delay_schedule(F, dpl)
    assert lv.size == T
    assert (int(lv[-1]), int(tv[-1])) == last
    assert np.bincount(lv, minlength=ml + 1).tolist() == counts


def test_oracle_schedule_f600(oracle):
    _, tv = oracle.delay_schedule(600, 8)
    assert tv.tolist() == F600


@pytest.mark.parametrize("F,dpl,ml,T,last,counts", KATS)
def test_library_schedule_matches(pkg, F, dpl, ml, T, last, counts):
    """The product's own schedule (xpcs_delay_schedule, pure host code in the C-ABI library)."""
    assert pkg.level_max(F, dpl) == ml
    lv, tv = pkg.delay_schedule(F, dpl)
    assert lv.size == T and (int(lv[-1]), int(tv[-1])) == last


def test_library_schedule_equals_oracle_sweep(pkg, oracle):
    for dpl in (2, 4, 8, 16):
        for F in list(range(2, 200)) + [255, 256, 257, 1023, 1024, 1025, 4097, 65536, 99999]:
            a = pkg.delay_schedule(F, dpl)
            b = oracle.delay_schedule(F, dpl)
            assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (F, dpl)

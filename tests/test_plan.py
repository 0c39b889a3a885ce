"""The planner (swb200.cu make_plan / estimate / estimate_chain) without a GPU: which kernel the library picks for the
BASELINE configurations, and that the CTA-chained engine (launch config 7) is only chosen where it can run."""
import pytest

from concurrentproject_b200 import api


def test_config2_runs_on_the_chained_engine():
    p = api.plan(100000, 100000)                      # BASELINE config 2, gap_init == gap_ext: linear-gap kernel
    assert (p["mode"], p["config"], p["rows"], p["two_sided"]) == (1, 7, 3, 1)
    # 3.1 ms measured on a B200 (profiles/r02_chain_experiments.txt): the estimate only steers the choice, but it should be close
    assert 5.0e6 < p["est_cycles"] < 7.5e6
    q = api.plan(100000, 100000, no_linear=True)
    assert (q["mode"], q["config"], q["two_sided"]) == (0, 7, 1)
    assert q["est_cycles"] > p["est_cycles"]


def test_chained_engine_only_where_it_can_run():
    # re-based lanes (scores beyond 32767), 32-bit lanes: the pair engine
    assert api.plan(4000000, 4000000, lanes=17)["config"] in (1, 2, 3)
    assert api.plan(1000000, 1000000, lanes=17)["config"] in (1, 2, 3)
    assert api.plan(100000, 100000, lanes=32)["config"] in (1, 2, 3)
    # more bands than 4 x 148 warps at the requested row count: asked for, not possible -> pair engine
    assert api.plan(1000000, 1000000, config=7, rows=2)["config"] != 7
    # several GPUs (a ring): never chained (the caller passes all SMs of the ring and no two-sided permission on old rings)
    assert api.plan(4000000, 4000000, lanes=17, sms=8 * 148)["config"] in (1, 2, 3)
    # switched off by request
    assert api.plan(100000, 100000, config=1)["config"] == 1
    api.configure("chain", 0)
    try:
        assert api.plan(100000, 100000)["config"] in (1, 2, 3)
    finally:
        api.configure("chain", 1)


@pytest.mark.parametrize("n", [5000, 20000, 50000, 150000, 250000])
def test_small_and_medium_pairs_pick_a_shape_that_fits(n):
    p = api.plan(n, n)
    assert p["config"] == 7 and p["rows"] in (1, 2, 3, 4, 6, 8)
    bands = -(-n // (64 * p["rows"]))
    half = (bands + 1) // 2 if p["two_sided"] else bands
    assert 2 * ((half + 3) // 4) <= 148 + 1 if p["two_sided"] else (bands + 3) // 4 <= 148

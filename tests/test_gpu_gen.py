"""The device-side seeded generator (csrc/swb_gen.cu) against its host mirror (concurrentproject_b200/rng.py), byte for
byte, and against the C copy in the oracle.  SURVEY.md 8d: one portable counter-based generator for host, device and
Python instead of the reference harness's unseeded rand() % 4 (TestFileWithGPU.cpp:25-36)."""
import numpy as np
import pytest

import oracle_lib as O
from concurrentproject_b200 import rng

pytestmark = pytest.mark.gpu


def test_random_stream_matches_host_and_oracle():
    import torch
    from concurrentproject_b200 import api
    for seed, stream, n in [(2, 0, 100000), (3, 1, 33), (7, 12345678901, 4097), (1, 5, 1)]:
        out = torch.zeros(n, dtype=torch.uint8, device="cuda")
        api.gen_random_device(0, seed, stream, n, out.data_ptr())
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert np.array_equal(got, rng.random_acgt(seed, stream, n))
        assert np.array_equal(got, O.random_acgt(seed, stream, n))


def test_config4_and_config5_recipes_match_host():
    import torch
    from concurrentproject_b200 import api
    first, npairs, rl, wl = 1000003, 300, 150, 1000          # a shard that does not start at pair 0
    reads = torch.zeros((npairs, rl), dtype=torch.uint8, device="cuda")
    wins = torch.zeros((npairs, wl), dtype=torch.uint8, device="cuda")
    api.gen_read_pairs_device(0, 4, first, npairs, rl, wl, reads.data_ptr(), wins.data_ptr())
    torch.cuda.synchronize()
    r_h, w_h = reads.cpu().numpy(), wins.cpu().numpy()
    for p in range(npairs):
        r, w = rng.read_pair(4, first + p, rl, wl)
        assert np.array_equal(r_h[p], r) and np.array_equal(w_h[p], w), p
    # planted reads really are similar to their window, random reads are not
    planted = [O.gotoh_rolling(r_h[p], w_h[p]) for p in range(0, 40) if (first + p) % 2 == 0]
    unrelated = [O.gotoh_rolling(r_h[p], w_h[p]) for p in range(0, 40) if (first + p) % 2 == 1]
    assert min(planted) > 90 and max(unrelated) < 40
    npairs, L = 12, 10000
    a = torch.zeros((npairs, L), dtype=torch.uint8, device="cuda")
    b = torch.zeros((npairs, L), dtype=torch.uint8, device="cuda")
    api.gen_long_pairs_device(0, 5, 77, npairs, L, a.data_ptr(), b.data_ptr())
    torch.cuda.synchronize()
    a_h, b_h = a.cpu().numpy(), b.cpu().numpy()
    for p in range(npairs):
        x, y = rng.long_pair(5, 77 + p, L)
        assert np.array_equal(a_h[p], x) and np.array_equal(b_h[p], y), p

"""FASTA input path on the GPU: ragged, length-bucketed batches against the CPU oracle."""
import numpy as np
import pytest

import oracle_lib as O
from concurrentproject_b200 import rng

pytestmark = pytest.mark.gpu


def _ragged(seed, npairs, max_read, max_win, with_n=False):
    r = np.random.default_rng(seed)
    s1, s2 = [], []
    for k in range(npairs):
        wl = int(r.integers(0, max_win + 1))
        w = rng.random_acgt(seed, 2 * k, wl)
        rl = int(r.integers(0, max_read + 1))
        if wl > 8 and k % 2 == 0:                       # a mutated piece of the window
            o = int(r.integers(0, wl - min(rl, wl) + 1))
            rd = rng.mutate(w[o:o + min(rl, wl)], seed, 2 * k + 1, 0.05, 0.02)
        else:
            rd = rng.random_acgt(seed, 2 * k + 1, rl)
        w, rd = bytes(w), bytes(rd)
        if with_n and k % 17 == 3 and len(rd) > 2:
            rd = rd[:1] + b"N" + rd[2:]
        if k % 5 == 0:
            w, rd = rd, w                               # the shorter one is not always first
        s1.append(rd); s2.append(w)
    return s1, s2


def test_score_pairs_ragged_lengths_match_oracle(tmp_path):
    from concurrentproject_b200 import fasta
    s1, s2 = _ragged(11, 3000, 300, 1200)
    fasta.write_fasta(str(tmp_path / "a.fa"), [f"r{k}" for k in range(len(s1))], s1)
    fasta.write_fasta(str(tmp_path / "b.fa"), [f"w{k}" for k in range(len(s2))], s2, width=77)
    a, b = fasta.read_fasta(str(tmp_path / "a.fa")), fasta.read_fasta(str(tmp_path / "b.fa"))
    want = O.gotoh_batch(s1, s2)
    for mb in (1, 64, 100000):                          # one bucket per class ... everything merged into one
        got = fasta.score_pairs(a, b, min_bucket=mb)
        assert np.array_equal(got, want), mb
    p = (2, -3, 4, 1)
    assert np.array_equal(fasta.score_pairs(s1[:500], s2[:500], p, min_bucket=32), O.gotoh_batch(s1[:500], s2[:500], p))


def test_score_pairs_routes_other_alphabets_and_long_pairs_to_the_pair_engine():
    from concurrentproject_b200 import fasta
    s1, s2 = _ragged(12, 400, 200, 600, with_n=True)
    s1 += [bytes(rng.random_acgt(12, 9001, 1500)), b"MKVLAAGIVGLLLAQWSHA", b""]
    s2 += [bytes(rng.random_acgt(12, 9002, 2500)), b"MKVLSAGIVALLLAQPSHA", b"ACGT"]
    want = np.array([O.gotoh_rolling(x, y) for x, y in zip(s1, s2)], dtype=np.int32)
    assert np.array_equal(fasta.score_pairs(s1, s2, min_bucket=16), want)


def test_search_one_query_against_a_database():
    from concurrentproject_b200 import fasta
    db = [bytes(rng.random_acgt(13, k, 50 + (k * 37) % 900)) for k in range(700)]
    q = rng.mutate(db[123][10:160], 13, 5000, 0.04, 0.01)
    want = np.array([O.gotoh_rolling(q, d) for d in db], dtype=np.int32)
    got = fasta.search(q, db, min_bucket=64)
    assert np.array_equal(got, want) and int(np.argmax(got)) == 123


def test_host_batch_call_pipelines_chunks():
    """swb200_score_batch cuts the batch into chunks (copy of chunk k+1 overlaps packing and scoring chunk k):
    force ~60 chunks with ragged lengths, a shared sequence (every pair reads the same seq1 bytes) and offsets that
    are not monotone."""
    from concurrentproject_b200 import api, fasta
    api.configure("batch_chunk_bytes", "20000")
    try:
        _pipelined_chunk_cases(api, fasta)
    finally:
        api.configure("batch_chunk_bytes", "0")


def _pipelined_chunk_cases(api, fasta):
    s1, s2 = _ragged(21, 1500, 250, 900)
    assert np.array_equal(api.score_batch(s1, s2), O.gotoh_batch(s1, s2))
    db = [bytes(rng.random_acgt(22, k, 30 + (k * 53) % 700)) for k in range(900)]
    q = bytes(rng.mutate(db[500][5:140], 22, 7000, 0.05, 0.02))
    want = np.array([O.gotoh_rolling(q, d) for d in db], dtype=np.int32)
    assert np.array_equal(fasta.search(q, db, min_bucket=100000), want)          # seq1 offsets all zero
    perm = np.random.default_rng(3).permutation(len(s1))                          # non-monotone offsets
    f1, o1, l1 = api._flatten(s1); f2, o2, l2 = api._flatten(s2)
    got = api.score_batch_flat(f1, o1[perm], l1[perm], f2, o2[perm], l2[perm])
    assert np.array_equal(got, O.gotoh_batch(s1, s2)[perm])
    lo, hi = -32, 31
    a = [bytes(rng.random_acgt(23, k, 400 + k % 50)) for k in range(300)]
    b = [bytes(rng.mutate(x, 23, 1000 + k, 0.05, 0.01)) for k, x in enumerate(a)]
    assert np.array_equal(api.score_banded_batch(a, b, lo, hi), O.gotoh_banded_batch(a, b, lo, hi))


def test_packed_host_batches():
    """swb200_score_batch_packed: the host batch arrives in the resident 2-bit format (swb200_pack_batch_host, format
    conversion only); ragged lengths, either sequence the shorter one, many small chunks through the copy/compute pipeline."""
    from concurrentproject_b200 import api
    s1, s2 = _ragged(41, 1200, 300, 1000)
    s1[3], s2[3] = s2[3], s1[3]                                  # a pair whose FIRST sequence is the longer one
    f1, o1, l1 = api._flatten(s1); f2, o2, l2 = api._flatten(s2)
    qw, qs, tw, ts, ql, tl = api.pack_batch_host(f1, o1, l1, f2, o2, l2)
    assert np.all(ql <= tl)
    want = O.gotoh_batch(s1, s2)
    assert np.array_equal(api.score_batch_packed(qw, qs, tw, ts, ql, tl), want)
    assert np.array_equal(api.score_batch_packed(qw, qs, tw, ts, ql, tl, (2, -3, 5, 1)), O.gotoh_batch(s1, s2, (2, -3, 5, 1)))
    api.configure("batch_chunk_bytes", "30000")
    try:
        assert np.array_equal(api.score_batch_packed(qw, qs, tw, ts, ql, tl), want)
    finally:
        api.configure("batch_chunk_bytes", "0")
    with pytest.raises(api.SwbError):
        api.pack_batch_host(np.frombuffer(b"ACGN", dtype=np.uint8), np.zeros(1, np.int64), np.array([4], np.int32),
                            np.frombuffer(b"ACGT", dtype=np.uint8), np.zeros(1, np.int64), np.array([4], np.int32))


def test_packed_host_banded_batches():
    """swb200_score_banded_batch_packed: banded scoring of a host batch that arrives in the 2-bit format (seq1 = columns and
    seq2 = rows keep their roles, either may be the longer one); ragged lengths from empty to 2 500, three bands, two
    parameter sets, many small chunks."""
    from concurrentproject_b200 import api
    r = np.random.default_rng(77)
    s1, s2 = [], []
    for k in range(150):
        a = rng.random_acgt(77, k, int(r.integers(0, 2500)))
        b = rng.mutate(a, 77, 1000 + k, 0.08, 0.02) if k % 3 else rng.random_acgt(77, 2000 + k, int(r.integers(0, 2500)))
        s1.append(bytes(a)); s2.append(bytes(b))
    f1, o1, l1 = api._flatten(s1); f2, o2, l2 = api._flatten(s2)
    w1, st1, w2, st2 = api.pack_banded_host(f1, o1, l1, f2, o2, l2)
    for lo in (-32, 0, -63):
        for p in (O.DEFAULT, (2, -3, 5, 1)):
            want = O.gotoh_banded_batch(s1, s2, lo, lo + 63, p)
            assert np.array_equal(api.score_banded_batch_packed(w1, st1, w2, st2, l1, l2, lo, lo + 63, p), want), (lo, p)
    assert np.array_equal(api.score_banded_batch_packed(w1, st1, w2, st2, l1, l2), api.score_banded_batch(s1, s2))
    api.configure("batch_chunk_bytes", "30000")
    try:
        assert np.array_equal(api.score_banded_batch_packed(w1, st1, w2, st2, l1, l2), O.gotoh_banded_batch(s1, s2, -32, 31))
    finally:
        api.configure("batch_chunk_bytes", "0")
    with pytest.raises(api.SwbError):
        api.score_banded_batch_packed(w1, st1, w2, st2, l1, l2, -10, 10)

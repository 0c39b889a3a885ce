"""swb200_align: the alignment itself (second half).  The reference is score-only (README.md:6), so
the contract is: score and span as the oracle says, and the emitted CIGAR, re-scored HERE (independently of the library's
own check) with the costs of main.cpp:28-33,57-58 over exactly that span, gives the same int."""
import re

import numpy as np
import pytest

import oracle_lib as O
from concurrentproject_b200 import rng

pytestmark = pytest.mark.gpu


def rescore(cigar, a, b, span, p):
    """Walks the CIGAR over seq1[j_start-1 .. j_end) x seq2[i_start-1 .. i_end); returns (score, consumed everything)."""
    ma, mi, gi, ge = p
    i0, j0, i1, j1 = span
    i, j, total = i0 - 1, j0 - 1, 0
    for cnt, op in re.findall(r"(\d+)([=XID])", cigar):
        cnt = int(cnt)
        if op in "=X":
            seg_a, seg_b = np.asarray(a[j:j + cnt]), np.asarray(b[i:i + cnt])
            assert len(seg_a) == cnt and len(seg_b) == cnt
            assert bool(np.all(seg_a == seg_b)) if op == "=" else bool(np.all(seg_a != seg_b)), (op, i, j)
            total += cnt * (ma if op == "=" else mi)
            i += cnt; j += cnt
        else:
            total -= gi + (cnt - 1) * ge
            if op == "I":
                j += cnt
            else:
                i += cnt
    assert re.fullmatch(r"(\d+[=XID])*", cigar)
    return total, (i, j) == (i1, j1)


def test_alignment_rescoring_and_span_against_oracle():
    from concurrentproject_b200 import api
    cases = []
    for k, (n, sub, indel, p) in enumerate([(300, 0.06, 0.03, O.DEFAULT), (2000, 0.10, 0.05, (2, -3, 5, 1)), (5000, 0.08, 0.04, (5, -4, 11, 1)),
                                            (1500, 0.15, 0.08, (1, -1, 0, 0)), (3000, 0.05, 0.02, (3, -2, 2, 2)), (700, 0.3, 0.1, (2, -1, 3, 1))]):
        a = rng.random_acgt(800 + k, 0, n)
        b = rng.mutate(a, 800 + k, 1, sub, indel)
        if k % 2:                                            # the alignment sits inside longer, unrelated flanks
            a = np.concatenate([rng.random_acgt(800 + k, 2, 333), a, rng.random_acgt(800 + k, 3, 100)])
            b = np.concatenate([rng.random_acgt(800 + k, 4, 77), b, rng.random_acgt(800 + k, 5, 512)])
        cases.append((a, b, p))
    for a, b, p in cases:
        score, span, cigar = api.align(a, b, p)
        want = O.gotoh_span(a, b, p)
        assert (score,) + span == want, (p, score, span, want)
        total, whole = rescore(cigar, a, b, span, p)
        assert whole and total == score, (p, total, score, cigar[:80])
        assert cigar[-1] == "=" and re.match(r"\d+=", cigar)  # an optimal local alignment starts and ends with a match
        assert score == api.score(a, b, p)


def test_alignment_edge_cases():
    from concurrentproject_b200 import api
    assert api.align(b"ACGT" * 50, b"ACGT" * 50) == (200, (1, 1, 200, 200), "200=")
    assert api.align(b"AAAA", b"CCCC") == (0, (0, 0, 0, 0), "")
    assert api.align(b"", b"ACGT") == (0, (0, 0, 0, 0), "")
    s, span, cigar = api.align(b"TTTTACGTACGTTTTT", b"GGACGTACGTGG")
    assert (s, cigar) == (8, "8=") and span == (3, 5, 10, 12)
    # one deleted stretch: a single gap, opened once
    a = rng.random_acgt(850, 0, 4000)
    b = np.concatenate([a[:2000], a[2007:]])
    p = (2, -3, 5, 1)                                       # affine: one gap of 7 beats any split (with 1/-1/1/1 it only ties)
    s, span, cigar = api.align(a, b, p)
    m = re.fullmatch(r"(\d+)=7I(\d+)=", cigar)              # the gap may slide over equal neighbouring bases
    assert s == 2 * 3993 - (5 + 6 * 1) and m and int(m.group(1)) + int(m.group(2)) == 3993 and span == (1, 1, 3993, 4000)
    assert abs(int(m.group(1)) - 2000) <= 3
    s, span, cigar = api.align(a, b)                        # linear gaps: same score as a single gap, whatever the split
    assert s == 3993 - 7 and rescore(cigar, a, b, span, O.DEFAULT) == (s, True)
    # more than four distinct symbols: byte-compare kernels in all three passes (the reference compares raw bytes, main.cpp:28-33)
    x, y = b"HELLOWORLDGATTACAHELLO", b"XXWORLDGATTTACAYY"
    s, span, cigar = api.align(x, y)
    assert s == O.gotoh_rolling(x, y) and rescore(cigar, np.frombuffer(x, np.uint8), np.frombuffer(y, np.uint8), span, O.DEFAULT) == (s, True)


def test_alignment_of_the_config2_pair():
    """Full size: the 100 000 x 100 000 pair of BASELINE config 2 (5 GB of traceback directions in HBM): score as pinned,
    span as swb200_score_span says, and the ~190 000-operation alignment re-scores to the score."""
    from concurrentproject_b200 import api
    a, b = rng.random_acgt(2, 0, 100000), rng.random_acgt(2, 1, 100000)
    score, span, cigar = api.align(a, b)
    assert score == 11446
    assert (score,) + span == api.score_span(a, b)
    assert rescore(cigar, a, b, span, O.DEFAULT) == (score, True)

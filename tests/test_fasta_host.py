"""Host logic of the FASTA input path: parser, ragged gather, bucket planner.  No GPU."""
import numpy as np
import pytest

from concurrentproject_b200 import fasta


def test_parse_basic_and_wrapping(tmp_path):
    text = b">one first record\nACGT\nacgt\n\n>two\r\nGG TT\r\n>empty\n>last\nAC*\n"
    rec = fasta.parse_fasta(text)
    assert rec.names == ["one first record", "two", "empty", "last"]
    assert [bytes(rec.seq(k)) for k in range(4)] == [b"ACGTACGT", b"GGTT", b"", b"AC"]
    assert rec.offsets.tolist() == [0, 8, 12, 12] and rec.lengths.tolist() == [8, 4, 0, 2]
    p = tmp_path / "x.fa"
    fasta.write_fasta(str(p), ["a", "b"], [b"A" * 130, b"CG"], width=60)
    back = fasta.read_fasta(str(p))
    assert back.names == ["a", "b"] and bytes(back.seq(0)) == b"A" * 130 and bytes(back.seq(1)) == b"CG"


def test_parse_no_trailing_newline_and_errors():
    rec = fasta.parse_fasta(b">x\nAC\nGT")
    assert bytes(rec.seq(0)) == b"ACGT"
    assert len(fasta.parse_fasta(b"")) == 0
    with pytest.raises(ValueError):
        fasta.parse_fasta(b"ACGT\n")
    with pytest.raises(ValueError):
        fasta.parse_fasta(b"ACGT\n>x\nAC\n")


def test_parse_random_roundtrip(tmp_path):
    rng = np.random.default_rng(5)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=int(n))) for n in rng.integers(0, 400, size=200)]
    p = tmp_path / "r.fa"
    fasta.write_fasta(str(p), [f"s{k}" for k in range(len(seqs))], seqs, width=int(rng.integers(1, 90)))
    rec = fasta.read_fasta(str(p))
    assert [bytes(rec.seq(k)) for k in range(len(rec))] == seqs


def test_gather_and_acgt_mask():
    rec = fasta.parse_fasta(b">a\nACGT\n>b\nNNA\n>c\n\n>d\nGGGGG\n")
    ids = np.array([3, 0, 2, 0])
    flat, off, lens = fasta._gather(rec.flat, rec.offsets, rec.lengths, ids)
    assert bytes(flat) == b"GGGGGACGTACGT" and off.tolist() == [0, 5, 9, 9] and lens.tolist() == [5, 4, 0, 4]
    assert fasta._record_is_acgt(rec.flat, rec.offsets, rec.lengths.astype(np.int64)).tolist() == [True, False, True, True]


def test_plan_buckets_partition_and_order():
    rng = np.random.default_rng(7)
    n = 5000
    l1 = rng.integers(0, 1500, size=n)
    l2 = rng.integers(0, 3000, size=n)
    ok = rng.random(n) > 0.05
    buckets = fasta.plan_buckets(l1, l2, ok, min_bucket=64)
    seen = np.concatenate([b.index for b in buckets])
    assert sorted(seen.tolist()) == list(range(n))                       # a partition of the pair ids
    short, long_ = np.minimum(l1, l2), np.maximum(l1, l2)
    for b in buckets:
        if b.kind == "batch":
            assert ok[b.index].all() and short[b.index].max() <= b.cap <= fasta.BATCH_MAX_SHORT
            assert (np.diff(long_[b.index]) <= 0).all()                  # longest first
        else:
            assert (~ok[b.index] | (short[b.index] > fasta.BATCH_MAX_SHORT)).all()
    # tiny classes are merged upward, never dropped
    few = fasta.plan_buckets(np.array([10, 500, 40]), np.array([20, 600, 50]), np.array([True] * 3), min_bucket=512)
    assert len(few) == 1 and few[0].kind == "batch" and few[0].cap == 1024 and sorted(few[0].index.tolist()) == [0, 1, 2]


def test_pairs_beyond_the_batch_kernels_score_range_go_to_the_single_pair_engine():
    """match * min(len) above the 16-bit batch limit (e.g. match 40 with 1 kb reads) must not reject the whole bucket:
    plan_buckets routes such pairs to the single-pair engine."""
    from concurrentproject_b200 import fasta
    l1 = np.array([100, 900, 1000, 50]); l2 = np.array([1000, 1000, 1000, 60])
    ok = np.ones(4, dtype=bool)
    b = fasta.plan_buckets(l1, l2, ok, min_bucket=1, match=40)
    single = [x for x in b if x.kind == "single"]
    assert single and sorted(single[0].index.tolist()) == [1, 2]          # 40*900, 40*1000 > 32766-40; 40*100, 40*50 fit
    assert all(i not in (1, 2) for x in b if x.kind == "batch" for i in x.index.tolist())
    assert not [x for x in fasta.plan_buckets(l1, l2, ok, min_bucket=1, match=1) if x.kind == "single"]

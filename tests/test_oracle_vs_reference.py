"""CPU: the oracle against the unmodified reference classes (oracle/_ref, built from /root/reference by oracle/Makefile).
Skipped when the reference library is neither prebuilt nor buildable (no /root/reference)."""
import numpy as np
import pytest

from kmc_testlib import Params, synth_bin, pack_superkmers, choose_lut_prefix_len


@pytest.mark.parametrize("k,both,cmin", [(31, True, 2), (31, False, 1), (28, True, 1), (28, False, 2), (55, True, 2), (55, False, 1),
                                         (17, True, 1), (32, True, 2), (64, True, 1), (70, True, 2), (33, True, 1), (128, True, 1)])
def test_bin_matches_reference(oracle, reference, k, both, cmin):
    p = Params(k=k, both_strands=both, cutoff_min=cmin, lut_prefix_len=choose_lut_prefix_len(k))
    b = synth_bin(k + cmin, k, 2500, genome_len=3000, err=0.02)
    o = oracle.process_bin(b, p)
    assert o.same_as(reference.process_bin(b, p))                      # RADULS, one sorter
    assert o.same_as(reference.process_bin(b, p, sort_kind=1))         # radix.h + CSmallSort (non-Intel hosts, kmc.h:1556-1560)
    assert o.same_as(reference.process_bin(b, p, n_sorters=4))         # in-bin threads: packs with gaps (kxmer_set.h:299-314)


def test_cutoffs_and_clamp_match_reference(oracle, reference):
    for cmin, cmax, cntmax in [(1, 10 ** 9, 255), (3, 9, 4), (2, 300, 65535), (1, 10 ** 9, 1)]:
        p = Params(k=31, cutoff_min=cmin, cutoff_max=cmax, counter_max=cntmax, lut_prefix_len=7)
        b = synth_bin(cmin * 7 + cmax % 13, 31, 3000, genome_len=500, err=0.005)
        assert oracle.process_bin(b, p).same_as(reference.process_bin(b, p))


def test_edge_bins_match_reference(oracle, reference):
    rng = np.random.default_rng(1)
    p = Params(k=31, cutoff_min=1, lut_prefix_len=7)
    for b in [pack_superkmers(31, [rng.integers(0, 4, 31)]),
              pack_superkmers(31, [rng.integers(0, 4, 31 + 255) for _ in range(40)]),
              pack_superkmers(31, [np.zeros(31 + 255, dtype=np.uint8) for _ in range(30)]),
              pack_superkmers(31, [np.tile(np.array([0, 3], dtype=np.uint8), 100)[:31 + 150] for _ in range(20)])]:
        assert oracle.process_bin(b, p).same_as(reference.process_bin(b, p))


def test_several_bins_many_sorters(oracle, reference):
    p = Params(k=31, cutoff_min=2, lut_prefix_len=7)
    bins = [synth_bin(50 + i, 31, n, genome_len=max(n, 400)) for i, n in enumerate([400, 0, 2500, 30, 1200])]
    res, _ = reference.process_bins(bins, p, n_sorters=3)
    for b, r in zip(bins, res):
        assert oracle.process_bin(b, p).same_as(r)


@pytest.mark.parametrize("words,key_bytes", [(1, 8), (1, 5), (2, 14), (2, 15), (3, 18), (4, 32)])
def test_sort_matches_raduls(oracle, reference, words, key_bytes):
    rng = np.random.default_rng(words * 100 + key_bytes)
    n = 20000
    raw = rng.integers(0, 256, size=(n, words * 8), dtype=np.uint8)
    raw[:, key_bytes:] = 0
    raw[n // 2:] = raw[rng.integers(0, n // 2, n - n // 2)]
    recs = raw.view(np.uint64).reshape(n, words)
    ref_sorted, _ = reference.sort(recs, key_bytes, n_threads=2)
    assert np.array_equal(oracle.sort(recs, key_bytes), ref_sorted)

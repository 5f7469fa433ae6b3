"""Generates the committed golden fixtures from the REFERENCE itself (run in the dev container, where
/root/reference exists and oracle/_ref has been built by `make -C oracle ref`).

  kats.json      the reference's own CLI known-answer tests restated for one bin:
                   * tests/kmc_CLI/data/single_read.fq, k=28, -ci1 -> "#Total no. of k-mers" == 70   (.github/workflows/main.yml:35-38)
                   * tests/kmc_CLI/data/issue-180/input.fa, k=5, -ci1 -> kmc_dump == pattern.dump   (.github/workflows/main.yml:48-52)
  bins_*.npz     small synthetic bins with the outputs of the unmodified CKmerBinSorter<SIZE> (oracle/ref/ref_harness.cpp):
                 payload bytes, LUT and the four counters, for several k / strand / cutoff settings.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from kmc_testlib import Params, Reference, synth_bin, pack_superkmers, choose_lut_prefix_len  # noqa: E402

REF = "/root/reference"


def kats():
    fq = open(os.path.join(REF, "tests/kmc_CLI/data/single_read.fq")).read().split("\n")
    read = fq[1].strip()
    fa = [l.strip() for l in open(os.path.join(REF, "tests/kmc_CLI/data/issue-180/input.fa")) if l.strip() and not l.startswith(">")]
    dump = [l.split() for l in open(os.path.join(REF, "tests/kmc_CLI/data/issue-180/pattern.dump")) if l.strip()]
    return [
        {"name": "single_read_k28", "k": 28, "cutoff_min": 1, "lut_prefix_len": 4, "reads": [read], "n_total": 70},
        {"name": "issue_180_palindrome_k5", "k": 5, "cutoff_min": 1, "lut_prefix_len": 1, "reads": fa, "n_total": sum(len(r) - 5 + 1 for r in fa),
         "dump": [[a, int(b)] for a, b in dump]},
    ]


CASES = [
    # name, k, both_strands, cutoff_min, cutoff_max, counter_max, n_super_kmers, genome_len, err
    ("k31_canon", 31, True, 2, 10 ** 9, 255, 1500, 2500, 0.02),
    ("k31_all_ci1", 31, False, 1, 10 ** 9, 255, 1200, 2500, 0.02),
    ("k31_cx_cs", 31, True, 2, 12, 5, 1500, 600, 0.01),
    ("k28_kxmer", 28, True, 1, 10 ** 9, 255, 1200, 2000, 0.02),
    ("k55_kxmer", 55, True, 2, 10 ** 9, 255, 1200, 2000, 0.02),
    ("k55_all", 55, False, 1, 10 ** 9, 65535, 800, 1500, 0.02),
    ("k17", 17, True, 1, 10 ** 9, 255, 1500, 900, 0.0),
    ("k32_maxx0", 32, True, 2, 10 ** 9, 255, 1200, 2000, 0.02),
    ("k64_two_words", 64, True, 1, 10 ** 9, 255, 800, 1500, 0.02),
    ("k70_three_words", 70, True, 2, 10 ** 9, 255, 800, 1500, 0.02),
    ("k128_four_words", 128, True, 1, 10 ** 9, 255, 500, 1500, 0.02),
    ("k31_counter1", 31, True, 1, 10 ** 9, 1, 600, 800, 0.02),
]


def main():
    R = Reference()
    json.dump(kats(), open(os.path.join(HERE, "kats.json"), "w"), indent=1)
    for name, k, both, cmin, cmax, cntmax, nsk, glen, err in CASES:
        p = Params(k=k, both_strands=both, cutoff_min=cmin, cutoff_max=cmax, counter_max=cntmax, lut_prefix_len=choose_lut_prefix_len(k))
        b = synth_bin(hash(name) % 1000 + 1 if False else sum(map(ord, name)), k, nsk, genome_len=glen, err=err)
        r = R.process_bin(b, p)
        nz = np.nonzero(r.lut)[0]
        np.savez_compressed(os.path.join(HERE, "bins_%s.npz" % name), data=b.data, n_rec=b.n_rec, pack_bytes=b.pack_bytes,
                            extras=b.extras, pack_first=b.pack_first,
                            params=np.array([k, int(both), cmin, cmax, cntmax, p.lut_prefix_len], dtype=np.int64),
                            payload=np.frombuffer(r.payload, dtype=np.uint8), lut_idx=nz.astype(np.int64), lut_val=r.lut[nz],
                            stats=np.array(r.stats, dtype=np.uint64))
        print(name, b.n_rec, len(r.payload), r.stats)
    # the two KATs through the reference classes as well (pins the harness against the reference's published answers)
    from kmc_testlib import bin_from_reads, decode_payload
    for kat in kats():
        p = Params(k=kat["k"], cutoff_min=kat["cutoff_min"], lut_prefix_len=kat["lut_prefix_len"])
        r = R.process_bin(bin_from_reads(kat["k"], kat["reads"]), p)
        assert r.stats[3] == kat["n_total"], (kat["name"], r.stats)
        if "dump" in kat:
            assert decode_payload(r.payload, r.lut, p) == [tuple(x) for x in kat["dump"]], kat["name"]
        print("KAT ok:", kat["name"])


if __name__ == "__main__":
    main()

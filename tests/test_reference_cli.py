"""BASELINE configs[0] (CPU plumbing, no GPU): a ~1 MB synthetic FASTQ through the UNMODIFIED reference CLI (oracle/_ref/kmc_ref, built by
`make -C oracle cli`) at k = 15 (the regular bin pipeline: the small-k direct-count path needs k <= 13) and k = 13 (small-k
path), checked against a brute-force count of the reads (the reference's own test strategy: tests/kmc_CLI/trivial-k-mer-counter).  This pins
the test infrastructure the whole-file GPU parity tests rest on (the FASTQ writer, the CLI wrappers, kmc_tools dump)."""
import os

import pytest

from test_gpu_kmc_files import KMC_REF, KMC_TOOLS, write_fastq, count, dump_sorted


def brute_force(fastq, k, both=True):
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    cnt = {}
    with open(fastq) as f:
        for i, line in enumerate(f):
            if i % 4 != 1:
                continue
            for piece in line.strip().split("N"):
                for j in range(len(piece) - k + 1):
                    km = piece[j:j + k]
                    if both:
                        rc = "".join(comp[c] for c in reversed(km))
                        km = min(km, rc)
                    cnt[km] = cnt.get(km, 0) + 1
    return cnt


@pytest.mark.parametrize("k", [15, 13])
def test_reference_cli_on_a_small_fastq(tmp_path, k):
    if not (os.path.exists(KMC_REF) and os.path.exists(KMC_TOOLS)):
        pytest.skip("oracle/_ref/kmc_ref not built (make -C oracle cli needs /root/reference)")
    tmp = str(tmp_path)
    fq = os.path.join(tmp, "reads.fq")
    write_fastq(fq, 15, 3400, genome_len=200_000)          # 3400 x 150 bp: ~1 MB of FASTQ
    assert 0.9e6 < os.path.getsize(fq) < 1.3e6
    db, stats = count(KMC_REF, tmp, "ref", fq, k, ("-ci2", "-cs255"))
    exp = {km: min(c, 255) for km, c in brute_force(fq, k).items() if c >= 2}
    got = dict(l.split() for l in dump_sorted(tmp, db, "ref").splitlines())
    assert {km: int(c) for km, c in got.items()} == exp
    st = stats["Stats"]
    assert int(st["#Unique_counted_k-mers"]) == len(exp)

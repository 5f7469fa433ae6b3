"""CPU: the C-ABI library builds, loads and exports every symbol that include/kmc_b200.h declares; host-only entry
points work; the product fails loudly (no CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "kmc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(kmcb200_[a-z_0-9]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    import kmc_b200
    L = kmc_b200.load_library()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), "missing export " + s
    assert sorted(kmc_b200.EXPORTS) == syms


def test_product_library_holds_no_test_code():
    """The synthetic-bin generator lives in tests/synth (the reference arm of bench.py must not map product code for it)."""
    import kmc_b200
    L = kmc_b200.load_library()
    assert not hasattr(L, "kmcb200_synth_bin") and not hasattr(L, "kmcsynth_bin")


def test_synth_generator_and_oracle_walk(oracle):
    from kmc_testlib import Bin, fast_bin
    sk = fast_bin(42, 31, 100000, genome_len=5000)
    assert sk.n_rec == 100000 and sk.pack_bytes.sum() == sk.size and np.all(sk.pack_bytes <= 1 << 16)
    b = Bin(data=sk.data, n_rec=sk.n_rec, n_super_kmers=sk.n_super_kmers, pack_bytes=sk.pack_bytes, pack_recs=sk.pack_bytes, k=31)
    assert oracle.walk(b) == (sk.n_super_kmers, sk.n_rec)
    # packs end on record boundaries
    pos = 0
    for pb in sk.pack_bytes[:5]:
        sub = Bin(data=sk.data[pos:pos + int(pb)], n_rec=0, n_super_kmers=0, pack_bytes=None, pack_recs=None, k=31)
        assert oracle.walk(sub)[0] > 0
        pos += int(pb)
    # deterministic
    sk2 = fast_bin(42, 31, 100000, genome_len=5000)
    assert np.array_equal(sk.data, sk2.data)


def test_create_fails_loudly_without_gpu():
    import torch
    import kmc_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(kmc_b200.KmcB200Error) as ei:
        kmc_b200.Stage2Context(kmc_b200.Stage2Params())
    assert ei.value.code == kmc_b200.ERR_NO_DEVICE
    assert "no CPU fallback" in str(ei.value)


def test_create_rejects_bad_params():
    import kmc_b200
    L = kmc_b200.load_library()
    h = C.c_void_p()
    for k, p in [(0, 3), (129, 5), (31, 6), (31, 0), (31, 31)]:
        prm = kmc_b200._Params(k, 1, 2, 1000, 255, p, 0, 1)
        assert L.kmcb200_create(C.byref(prm), C.byref(h)) == kmc_b200.ERR_INVALID
        assert len(L.kmcb200_last_error(None)) > 0

"""kmc_b200 — Python host side over the C ABI (include/kmc_b200.h) of the B200 stage-2 path of KMC.

The product is the CUDA library `libkmc_b200.so` (kmc_b200/csrc); this module is a thin ctypes mirror used
by the tests and by bench.py.  Names follow the reference: a *bin* of super-k-mers goes through
Expand -> Sort -> Compact exactly like `CKmerBinSorter<SIZE>::ProcessBins` (kmc_core/kb_sorter.h:210-237).
There is no CPU fallback: constructing a `Stage2Context` without a B200 raises.
"""
import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.abspath(os.environ["KMCB200_LIB"]) if os.environ.get("KMCB200_LIB") else os.path.join(_HERE, "libkmc_b200.so")      # KMCB200_LIB: experiment builds (kmc_b200.build.build_variant)

OK, ERR_INVALID, ERR_NO_DEVICE, ERR_CUDA, ERR_BIN_FORMAT, ERR_CAPACITY, ERR_BUSY = 0, -1, -2, -3, -4, -5, -6


class KmcB200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("kmc_b200 error %d: %s" % (code, msg))
        self.code = code


class _Params(C.Structure):
    _fields_ = [("kmer_len", C.c_uint32), ("both_strands", C.c_uint32), ("cutoff_min", C.c_uint32), ("cutoff_max", C.c_uint32),
                ("counter_max", C.c_uint32), ("lut_prefix_len", C.c_uint32), ("device", C.c_int32), ("n_slots", C.c_uint32)]


class _DbParams(C.Structure):
    _fields_ = [("kmer_len", C.c_uint32), ("counter_size", C.c_uint32), ("lut_prefix_len", C.c_uint32), ("signature_len", C.c_uint32),
                ("cutoff_min", C.c_uint32), ("cutoff_max", C.c_uint32), ("both_strands", C.c_uint32)]


EXPORTS = [
    "kmcb200_create", "kmcb200_destroy", "kmcb200_last_error", "kmcb200_out_rec_bytes", "kmcb200_out_capacity", "kmcb200_lut_entries",
    "kmcb200_host_alloc", "kmcb200_host_free", "kmcb200_process_bin", "kmcb200_process_bin_multi", "kmcb200_submit_bin", "kmcb200_submit_bin_indexed", "kmcb200_wait_bin", "kmcb200_sort_records",
    "kmcb200_dev_process_bin", "kmcb200_dev_expand", "kmcb200_dev_sort", "kmcb200_dev_count", "kmcb200_kernel_launches",
    "kmcb200_stage_times", "kmcb200_stage_names",
    "kmcb200_wait_bin_scanned", "kmcb200_db_open", "kmcb200_db_last_error", "kmcb200_db_records", "kmcb200_db_reserve", "kmcb200_db_commit_bin", "kmcb200_db_close",
]

_lib = None


def load_library(build_if_needed=True):
    """Loads (building it first when nvcc is around and the sources are newer) the CUDA library. Fails loudly if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if build_if_needed and os.path.exists(_build.NVCC) and _build.needs_build():
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise KmcB200Error(ERR_NO_DEVICE, "%s is missing: run `python -m kmc_b200.build` (there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32
    L.kmcb200_create.argtypes = [C.POINTER(_Params), C.POINTER(vp)]
    L.kmcb200_destroy.argtypes = [vp]
    L.kmcb200_destroy.restype = None
    L.kmcb200_last_error.argtypes = [vp]
    L.kmcb200_last_error.restype = C.c_char_p
    L.kmcb200_out_rec_bytes.argtypes = [vp]
    L.kmcb200_out_rec_bytes.restype = u32
    L.kmcb200_out_capacity.argtypes = [vp, u64]
    L.kmcb200_out_capacity.restype = u64
    L.kmcb200_lut_entries.argtypes = [vp]
    L.kmcb200_lut_entries.restype = u64
    L.kmcb200_host_alloc.argtypes = [vp, u64, C.POINTER(vp)]
    L.kmcb200_host_free.argtypes = [vp, vp]
    L.kmcb200_process_bin.argtypes = [vp, i32, vp, u64, u64, u64, vp, vp, u32, vp, u64, C.POINTER(u64), vp, vp]
    L.kmcb200_process_bin_multi.argtypes = [vp, u32, i32, vp, u64, u64, vp, u32, vp, u64, C.POINTER(u64), vp, vp]
    L.kmcb200_submit_bin.argtypes = [vp, u32, i32, vp, u64, u64, u64, vp, vp, u32, vp, u64, vp]
    L.kmcb200_submit_bin_indexed.argtypes = [vp, u32, i32, vp, u64, u64, vp, u32, vp, u64, vp, vp, u64, vp]
    L.kmcb200_wait_bin.argtypes = [vp, u32, C.POINTER(u64), vp]
    L.kmcb200_sort_records.argtypes = [vp, vp, vp, u64, u32, u32]
    L.kmcb200_dev_process_bin.argtypes = [vp, u32, vp, u64, u64, vp, u32, vp, u64, vp, vp, vp]
    L.kmcb200_dev_expand.argtypes = [vp, u32, vp, u64, u64, vp, u32, vp, vp, vp]
    L.kmcb200_dev_sort.argtypes = [vp, u32, vp, vp, u64, u32, C.c_int, vp]
    L.kmcb200_dev_count.argtypes = [vp, u32, vp, u64, vp, u64, vp, vp, vp]
    L.kmcb200_kernel_launches.argtypes = [vp]
    L.kmcb200_kernel_launches.restype = u64
    L.kmcb200_stage_times.argtypes = [vp, u32, C.POINTER(C.c_float), u32]
    L.kmcb200_stage_names.argtypes = [vp, u32, C.c_char_p, u32]
    L.kmcb200_wait_bin_scanned.argtypes = [vp, u32, u64, C.POINTER(u64), vp]
    L.kmcb200_db_open.argtypes = [C.POINTER(_DbParams), C.c_char_p, u64, C.POINTER(vp)]
    L.kmcb200_db_last_error.argtypes = [vp]
    L.kmcb200_db_last_error.restype = C.c_char_p
    L.kmcb200_db_records.argtypes = [vp]
    L.kmcb200_db_records.restype = u64
    L.kmcb200_db_reserve.argtypes = [vp, u64, C.POINTER(vp)]
    L.kmcb200_db_commit_bin.argtypes = [vp, u64, vp, C.c_int, vp, vp, u32]
    L.kmcb200_db_close.argtypes = [vp, vp]
    _lib = L
    return L


@dataclass
class Stage2Params:
    """The per-run parameters CKmerBinSorter takes from CKMCParams (kmc_core/kb_sorter.h:165-200)."""
    kmer_len: int = 31
    both_strands: bool = True
    cutoff_min: int = 2
    cutoff_max: int = 1_000_000_000
    counter_max: int = 255
    lut_prefix_len: int = 7


@dataclass
class SuperKmerBin:
    """One bin as stage 1 leaves it: byte stream + CBinDesc counters + expander packs (queues.h:376-679)."""
    data: np.ndarray            # uint8
    n_rec: int
    pack_bytes: np.ndarray      # uint64
    n_super_kmers: int = 0
    kmer_len: int = 31

    @property
    def size(self):
        return int(self.data.size)


@dataclass
class BinResult:
    """What CKmerBinSorter hands to CKmerQueue::push (queues.h:826): emitted records, raw LUT, the four counters."""
    payload: np.ndarray         # uint8, concatenated (suffix, counter) records
    lut: np.ndarray             # uint64[4^p]
    n_unique: int
    n_cutoff_min: int
    n_cutoff_max: int
    n_total: int

    @property
    def stats(self):
        return (self.n_unique, self.n_cutoff_min, self.n_cutoff_max, self.n_total)


class Stage2Context:
    """One GPU's stage-2 engine (one per sorter thread in KMC terms)."""

    def __init__(self, params: Stage2Params, device=0, n_slots=1):
        self.lib = load_library()
        self.params = params
        self._h = C.c_void_p(None)
        p = _Params(params.kmer_len, int(params.both_strands), params.cutoff_min, min(params.cutoff_max, 0xFFFFFFFF),
                    min(params.counter_max, 0xFFFFFFFF), params.lut_prefix_len, device, n_slots)
        rc = self.lib.kmcb200_create(C.byref(p), C.byref(self._h))
        if rc != 0:
            raise KmcB200Error(rc, (self.lib.kmcb200_last_error(None) or b"").decode())
        self.device = device
        self.n_slots = n_slots
        self.words = (params.kmer_len + 31) // 32
        self.key_bytes = (params.kmer_len + 3) // 4
        self.out_rec_bytes = self.lib.kmcb200_out_rec_bytes(self._h)
        self.lut_entries = self.lib.kmcb200_lut_entries(self._h)

    # -- plumbing
    def _check(self, rc):
        if rc < 0:
            raise KmcB200Error(rc, (self.lib.kmcb200_last_error(self._h) or b"").decode())
        return rc

    def close(self):
        if self._h:
            self.lib.kmcb200_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def out_capacity(self, n_rec):
        return int(self.lib.kmcb200_out_capacity(self._h, n_rec))

    def kernel_launches(self):
        return int(self.lib.kmcb200_kernel_launches(self._h))

    def stage_times(self, slot=0):
        ms = (C.c_float * 48)()
        n = self._check(self.lib.kmcb200_stage_times(self._h, slot, ms, 48))
        names = C.create_string_buffer(2048)
        self._check(self.lib.kmcb200_stage_names(self._h, slot, names, 2048))
        nm = names.value.decode().split(",") if names.value else []
        return {"expand_ms": ms[0], "sort_ms": ms[1], "count_ms": ms[2], "pass_ms": [ms[3 + i] for i in range(n)], "pass_names": nm}

    # -- seam #2, host buffers
    def process_bin(self, b: SuperKmerBin, out=None, lut=None, bin_id=0) -> BinResult:
        cap = self.out_capacity(b.n_rec) + 64
        out = np.empty(cap, dtype=np.uint8) if out is None else out
        lut = np.empty(self.lut_entries, dtype=np.uint64) if lut is None else lut
        stats = (C.c_uint64 * 4)()
        nbytes = C.c_uint64(0)
        data = np.ascontiguousarray(b.data)
        packs = np.ascontiguousarray(b.pack_bytes, dtype=np.uint64)
        self._check(self.lib.kmcb200_process_bin(self._h, bin_id, data.ctypes.data, data.size, b.n_rec, b.n_rec, packs.ctypes.data, None, packs.size,
                                                 out.ctypes.data, out.size, C.byref(nbytes), lut.ctypes.data, stats))
        return BinResult(out[:nbytes.value], lut, *[int(x) for x in stats])

    def submit_bin(self, slot, data_ptr, size, n_rec, packs: np.ndarray, out_ptr, out_capacity, lut_ptr, bin_id=0):
        self._check(self.lib.kmcb200_submit_bin(self._h, slot, bin_id, data_ptr, size, n_rec, n_rec, packs.ctypes.data, None, packs.size,
                                                out_ptr, out_capacity, lut_ptr))

    def submit_bin_indexed(self, slot, data_ptr, size, n_rec, packs: np.ndarray, extras: np.ndarray, pack_superkmers: np.ndarray, out_ptr, out_capacity, lut_ptr, bin_id=0):
        """submit_bin with stage 1's length bytes as a separate array (N4: no record walk on the GPU)."""
        extras = np.ascontiguousarray(extras, dtype=np.uint8)
        psk = np.ascontiguousarray(pack_superkmers, dtype=np.uint32)
        self._keep = (extras, psk)
        self._check(self.lib.kmcb200_submit_bin_indexed(self._h, slot, bin_id, data_ptr, size, n_rec, packs.ctypes.data, packs.size,
                                                        extras.ctypes.data, extras.size, psk.ctypes.data, out_ptr, out_capacity, lut_ptr))

    def wait_bin_scanned(self, slot, lut_base):
        """wait_bin with the LUT already prefix-summed on the GPU and offset by lut_base (what goes into .kmc_pre)."""
        stats = (C.c_uint64 * 4)()
        nbytes = C.c_uint64(0)
        self._check(self.lib.kmcb200_wait_bin_scanned(self._h, slot, lut_base, C.byref(nbytes), stats))
        return int(nbytes.value), tuple(int(x) for x in stats)

    def wait_bin(self, slot):
        stats = (C.c_uint64 * 4)()
        nbytes = C.c_uint64(0)
        self._check(self.lib.kmcb200_wait_bin(self._h, slot, C.byref(nbytes), stats))
        return int(nbytes.value), tuple(int(x) for x in stats)

    @staticmethod
    def process_bin_multi(ctxs, b: "SuperKmerBin", out=None, lut=None) -> "BinResult":
        """One bin split over several contexts / GPUs by key range (kmcb200_process_bin_multi)."""
        c0 = ctxs[0]
        cap = c0.out_capacity(b.n_rec) + 64
        out = np.empty(cap, dtype=np.uint8) if out is None else out
        lut = np.empty(c0.lut_entries, dtype=np.uint64) if lut is None else lut
        stats = (C.c_uint64 * 4)()
        nbytes = C.c_uint64(0)
        data = np.ascontiguousarray(b.data)
        packs = np.ascontiguousarray(b.pack_bytes, dtype=np.uint64)
        arr = (C.c_void_p * len(ctxs))(*[c._h for c in ctxs])
        c0._check(c0.lib.kmcb200_process_bin_multi(arr, len(ctxs), 0, data.ctypes.data, data.size, b.n_rec, packs.ctypes.data, packs.size,
                                                   out.ctypes.data, out.size, C.byref(nbytes), lut.ctypes.data, stats))
        return BinResult(out[:nbytes.value], lut, *[int(x) for x in stats])

    # -- seam #1
    def sort_records(self, recs: np.ndarray, key_bytes=None):
        """recs: uint64 [n, words].  Returns the sorted copy (SortFunction contract, raduls.h:19-20)."""
        recs = np.ascontiguousarray(recs, dtype=np.uint64).copy()
        n, w = recs.shape
        tmp = np.empty_like(recs)
        kb = self.key_bytes if key_bytes is None else key_bytes
        where = self._check(self.lib.kmcb200_sort_records(self._h, recs.ctypes.data, tmp.ctypes.data, n, 8 * w, kb))
        return tmp if where == 1 else recs

    # -- device-level (pointers are device pointers: ints or torch tensors' data_ptr())
    def dev_process_bin(self, slot, d_bin, size, n_rec, packs: np.ndarray, d_out, out_capacity, d_lut, d_result, stream=None):
        packs = np.ascontiguousarray(packs, dtype=np.uint64)
        self._check(self.lib.kmcb200_dev_process_bin(self._h, slot, d_bin, size, n_rec, packs.ctypes.data, packs.size, d_out, out_capacity, d_lut, d_result, stream))

    def dev_expand(self, slot, d_bin, size, n_rec, packs: np.ndarray, d_recs, d_result=None, stream=None):
        packs = np.ascontiguousarray(packs, dtype=np.uint64)
        self._check(self.lib.kmcb200_dev_expand(self._h, slot, d_bin, size, n_rec, packs.ctypes.data, packs.size, d_recs, d_result, stream))

    def dev_sort(self, slot, d_recs, d_tmp, n, key_bytes=None, hist_ready=False, stream=None):
        return self._check(self.lib.kmcb200_dev_sort(self._h, slot, d_recs, d_tmp, n, self.key_bytes if key_bytes is None else key_bytes, int(hist_ready), stream))

    def dev_count(self, slot, d_sorted, n, d_out, out_capacity, d_lut, d_result, stream=None):
        self._check(self.lib.kmcb200_dev_count(self._h, slot, d_sorted, n, d_out, out_capacity, d_lut, d_result, stream))


class DbWriter:
    """KMC database files from per-bin results (kmcb200_db_*): pinned staging ring + writer thread; format of kb_completer.cpp:59-326."""

    def __init__(self, path_prefix, kmer_len, counter_size, lut_prefix_len, signature_len, cutoff_min, cutoff_max, both_strands, staging_bytes=1 << 28):
        self.lib = load_library()
        self._h = C.c_void_p(None)
        p = _DbParams(kmer_len, counter_size, lut_prefix_len, signature_len, cutoff_min, min(cutoff_max, 0xFFFFFFFF), int(both_strands))
        rc = self.lib.kmcb200_db_open(C.byref(p), path_prefix.encode(), staging_bytes, C.byref(self._h))
        if rc != 0:
            raise KmcB200Error(rc, (self.lib.kmcb200_db_last_error(None) or b"").decode())
        self.lut_entries = 1 << (2 * lut_prefix_len)

    def _check(self, rc):
        if rc != 0:
            raise KmcB200Error(rc, (self.lib.kmcb200_db_last_error(self._h) or b"").decode())

    @property
    def records(self):
        return int(self.lib.kmcb200_db_records(self._h))

    def reserve(self, nbytes):
        ptr = C.c_void_p(None)
        self._check(self.lib.kmcb200_db_reserve(self._h, nbytes, C.byref(ptr)))
        return ptr.value

    def commit_bin(self, payload_bytes, lut: np.ndarray, stats, signatures=(), raw_lut=False):
        lut = np.ascontiguousarray(lut, dtype=np.uint64)
        assert lut.size == self.lut_entries
        st = (C.c_uint64 * 4)(*[int(x) for x in stats])
        sig = np.ascontiguousarray(np.asarray(signatures, dtype=np.uint32))
        self._check(self.lib.kmcb200_db_commit_bin(self._h, payload_bytes, lut.ctypes.data, int(raw_lut), st, sig.ctypes.data if sig.size else None, sig.size))

    def close(self):
        tot = (C.c_uint64 * 4)()
        h, self._h = self._h, C.c_void_p(None)
        rc = self.lib.kmcb200_db_close(h, tot)
        if rc != 0:
            raise KmcB200Error(rc, "kmcb200_db_close")
        return tuple(int(x) for x in tot)

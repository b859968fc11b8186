"""
ctypes binding of libtrtools_b200.so (include/trtools_b200.h).

There is no CPU fallback: if the library is missing, or no CUDA device is present, every
compute entry point raises.  ``load()`` only dlopens the library (no CUDA call), so the
symbol-export test can run on a box without a GPU.
"""
import ctypes as C
import os
from typing import Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TRTOOLS_B200_LIB", os.path.join(HERE, "libtrtools_b200.so"))

TRT_OK = 0
TRT_ENODEV, TRT_ECUDA, TRT_EINVAL, TRT_ESTATE, TRT_ENOMEM, TRT_ERECORD, TRT_ENCCL = -1, -2, -3, -4, -5, -6, -7

VCF_TYPES = {"gangstr": 0, "advntr": 1, "hipstr": 2, "eh": 3, "popstr": 4, "longtr": 5}
FMT_DP, FMT_DSTUTTER, FMT_DFLANKINDEL, FMT_Q, FMT_QEXP, FMT_AUX0 = range(6)
FMT_NAUX = 8
CF_MIN, CF_MAX, CF_RATIO_GT, CF_QEXP_HET, CF_QEXP_HOM, CF_QEXP_TOT, CF_HOST_VALUE = range(7)
LF_CALLRATE, LF_HWE, LF_HETLOW, LF_HETHIGH, LF_HRUN = range(5)
AF_OK, AF_NO_CALLED, AF_ONE_ALLELE, AF_NON_MAJOR, AF_NCOVARS = range(5)
HF_HAS_FULL, HF_MOTIF_N, HF_MOTIF_NONACGT, HF_LEN_DUPS, HF_SEQ_DUPS, HF_BAD_PERIOD = 1, 2, 4, 8, 16, 32

# every symbol include/trtools_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "trt_device_count", "trt_init", "trt_destroy", "trt_last_error", "trt_device_info", "trt_synchronize",
    "trt_host_alloc", "trt_host_free", "trt_launch_count", "trt_last_kernel_ms", "trt_last_scan_ms",
    "trt_stopwatch_start", "trt_stopwatch_stop",
    "trt_block_begin", "trt_block_set_gt", "trt_block_set_gt_packed", "trt_block_get_gt_packed", "trt_block_set_gt_nibble", "trt_block_get_gt_nibble", "trt_block_set_gt_device", "trt_block_set_format_i32",
    "trt_block_set_format_f32", "trt_block_set_format_device", "trt_block_set_alleles",
    "trt_harmonize", "trt_get_harmonized", "trt_pack_length_genotypes", "trt_get_packed_gt",
    "trt_locus_stats", "trt_genotype_counts", "trt_call_filters", "trt_locus_filters", "trt_assoc_set_design", "trt_assoc_ols",
    "trt_block_set_ap", "trt_dosages", "trt_assoc_dosage_ols", "trt_qc_reduce", "trt_compare",
    "trt_synth_fill", "trt_block_get_gt", "trt_block_get_format",
    "trt_dist_unique_id", "trt_dist_init", "trt_dist_allgather_f64", "trt_dist_allreduce_sum_i64",
    "trt_dist_allreduce_sum_f64", "trt_dist_allreduce_max_f64", "trt_dist_barrier", "trt_dist_gather_region",
    "trt_dist_gather_host", "trt_dist_wait", "trt_dist_finalize",
    "trt_vcf_open", "trt_vcf_close", "trt_vcf_last_error", "trt_vcf_header", "trt_vcf_n_samples",
    "trt_vcf_set_samples", "trt_vcf_seek", "trt_vcf_read_block", "trt_vcf_block_free", "trt_vcf_block_text", "trt_vcf_block_parse", "trt_vcf_block_parse_packed", "trt_vcf_block_parse_nibble", "trt_vcf_block_field", "trt_vcf_join_samples",
]


class TrtError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("trtools_b200 [{}]: {}".format(code, msg))
        self.code = code
        self.msg = msg


class DevInfo(C.Structure):
    _fields_ = [("name", C.c_char * 128), ("cc_major", C.c_int32), ("cc_minor", C.c_int32),
                ("sm_count", C.c_int32), ("total_mem_bytes", C.c_int64), ("free_mem_bytes", C.c_int64),
                ("l2_bytes", C.c_int32), ("abi_version", C.c_int32)]


class HarmonizeOut(C.Structure):
    _fields_ = [("allele_len", C.c_void_p), ("trim_off", C.c_void_p), ("trim_len", C.c_void_p),
                ("len_class", C.c_void_p), ("seq_class", C.c_void_p), ("len_order", C.c_void_p),
                ("seq_order", C.c_void_p), ("hrun", C.c_void_p), ("flags", C.c_void_p),
                ("motif", C.c_void_p), ("motif_off", C.c_void_p)]


class LocusStatsOut(C.Structure):
    _fields_ = [("ac", C.c_void_p), ("n_called", C.c_void_p), ("n_called_nonstrict", C.c_void_p),
                ("n_hom", C.c_void_p), ("n_padded", C.c_void_p), ("thresh", C.c_void_p), ("het", C.c_void_p),
                ("entropy", C.c_void_p), ("mean", C.c_void_p), ("mode", C.c_void_p), ("var", C.c_void_p),
                ("hwep", C.c_void_p), ("nalleles", C.c_void_p)]


class CallFilterSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("field_id", C.c_int32), ("threshold", C.c_double)]


class CallFilterOut(C.Structure):
    _fields_ = [("call_mask", C.c_void_p), ("trigger_values", C.c_void_p), ("gt_masked", C.c_void_p),
                ("filter_counts", C.c_void_p), ("numcalls", C.c_void_p), ("totaldp", C.c_void_p),
                ("negative_dp_locus", C.c_void_p)]


class LocusFilterSpec(C.Structure):
    _fields_ = [("kind", C.c_int32), ("threshold", C.c_double)]


class LocusFilterOut(C.Structure):
    _fields_ = [("flags", C.c_void_p), ("n_called", C.c_void_p), ("het", C.c_void_p), ("hwep", C.c_void_p),
                ("ac", C.c_void_p), ("hrun", C.c_void_p)]


class AssocOut(C.Structure):
    _fields_ = [("filter_code", C.c_void_p), ("n_tested", C.c_void_p), ("p", C.c_void_p), ("coef", C.c_void_p),
                ("se", C.c_void_p), ("r2", C.c_void_p), ("std_g", C.c_void_p), ("ac_len", C.c_void_p)]


class AssocDosageOut(C.Structure):
    _fields_ = [("n_tested", C.c_void_p), ("p", C.c_void_p), ("coef", C.c_void_p), ("se", C.c_void_p), ("r2", C.c_void_p),
                ("std_g", C.c_void_p), ("ncovars_code", C.c_void_p), ("class_stats", C.c_void_p), ("length_stats", C.c_void_p)]


class QcOut(C.Structure):
    _fields_ = [("sample_calls", C.c_void_p), ("locus_calls", C.c_void_p), ("sample_quality", C.c_void_p),
                ("locus_quality", C.c_void_p)]


class CompareIn(C.Structure):
    _fields_ = [("gt2", C.c_void_p), ("S2", C.c_int64), ("idx1", C.c_void_p), ("idx2", C.c_void_p), ("n_shared", C.c_int64),
                ("locus_off2", C.c_void_p), ("seq_id2", C.c_void_p), ("len2", C.c_void_p), ("reflen", C.c_void_p),
                ("ignore_phasing", C.c_int32)]


class CompareOut(C.Structure):
    _fields_ = [("numcalls", C.c_void_p), ("conc_seq", C.c_void_p), ("conc_len", C.c_void_p), ("len_sums", C.c_void_p),
                ("status", C.c_void_p), ("sample_numcalls", C.c_void_p), ("sample_conc_seq", C.c_void_p),
                ("sample_conc_len", C.c_void_p)]


DOSAGE_TYPES = {"bestguess": 0, "beagleap": 1, "bestguess_norm": 2, "beagleap_norm": 3}
DE_OK, DE_NO_AP, DE_AP_SUM, DE_AP_NEGATIVE, DE_NORM_RANGE = range(5)

_lib = None


def load():
    """dlopen the library and declare signatures (no CUDA call is made)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "libtrtools_b200.so is not built ({}). Run `python -m trtools_b200.build` "
            "(needs nvcc); there is no CPU fallback.".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, u32, u64, f64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64, C.c_double, C.c_size_t
    sig = {
        "trt_device_count": (i32, []),
        "trt_init": (i32, [i32, C.POINTER(vp)]),
        "trt_destroy": (None, [vp]),
        "trt_last_error": (C.c_char_p, [vp]),
        "trt_device_info": (i32, [vp, C.POINTER(DevInfo)]),
        "trt_synchronize": (i32, [vp]),
        "trt_host_alloc": (vp, [vp, sz]),
        "trt_host_free": (i32, [vp, vp]),
        "trt_launch_count": (i64, [vp]),
        "trt_last_kernel_ms": (f64, [vp]),
        "trt_last_scan_ms": (f64, [vp]),
        "trt_stopwatch_start": (i32, [vp]),
        "trt_stopwatch_stop": (i32, [vp, C.POINTER(f64)]),
        "trt_block_begin": (i32, [vp, i64, i64, i32, i32]),
        "trt_block_set_gt": (i32, [vp, vp]),
        "trt_block_set_gt_packed": (i32, [vp, vp, vp]),
        "trt_block_get_gt_packed": (i32, [vp, i64, i64, vp, vp]),
        "trt_block_set_gt_nibble": (i32, [vp, vp, vp]),
        "trt_block_get_gt_nibble": (i32, [vp, i64, i64, vp, vp]),
        "trt_block_set_gt_device": (i32, [vp, vp, sz]),
        "trt_block_set_format_i32": (i32, [vp, i32, vp]),
        "trt_block_set_format_f32": (i32, [vp, i32, vp, i32]),
        "trt_block_set_format_device": (i32, [vp, i32, vp, i32, i32]),
        "trt_block_set_alleles": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
        "trt_harmonize": (i32, [vp]),
        "trt_get_harmonized": (i32, [vp, C.POINTER(HarmonizeOut)]),
        "trt_pack_length_genotypes": (i32, [vp]),
        "trt_get_packed_gt": (i32, [vp, vp]),
        "trt_locus_stats": (i32, [vp, i32, vp, i32, f64, C.POINTER(LocusStatsOut)]),
        "trt_genotype_counts": (i32, [vp, i64, vp, vp, i64]),
        "trt_call_filters": (i32, [vp, C.POINTER(CallFilterSpec), i32, i32, C.POINTER(CallFilterOut)]),
        "trt_locus_filters": (i32, [vp, C.POINTER(LocusFilterSpec), i32, i32, C.POINTER(LocusFilterOut)]),
        "trt_assoc_set_design": (i32, [vp, vp, vp, vp, i64, i32]),
        "trt_assoc_ols": (i32, [vp, f64, C.POINTER(AssocOut)]),
        "trt_block_set_ap": (i32, [vp, vp, vp, vp]),
        "trt_dosages": (i32, [vp, i32, vp, vp]),
        "trt_assoc_dosage_ols": (i32, [vp, vp, vp, vp, C.POINTER(AssocDosageOut)]),
        "trt_qc_reduce": (i32, [vp, vp, vp, i32, i32, C.POINTER(QcOut)]),
        "trt_compare": (i32, [vp, C.POINTER(CompareIn), C.POINTER(CompareOut)]),
        "trt_synth_fill": (i32, [vp, u64, i64, i64, i64, vp, u32, u32, i32]),
        "trt_block_get_gt": (i32, [vp, i64, i64, vp]),
        "trt_block_get_format": (i32, [vp, i32, i64, i64, vp]),
        "trt_dist_unique_id": (i32, [vp]),
        "trt_dist_init": (i32, [vp, i32, i32, vp]),
        "trt_dist_allgather_f64": (i32, [vp, vp, i64, vp]),
        "trt_dist_allreduce_sum_i64": (i32, [vp, vp, i64]),
        "trt_dist_allreduce_sum_f64": (i32, [vp, vp, i64]),
        "trt_dist_allreduce_max_f64": (i32, [vp, vp, i64]),
        "trt_dist_barrier": (i32, [vp]),
        "trt_dist_gather_region": (i32, [vp, i32, i64, i64, vp, i32, vp, i32]),
        "trt_dist_gather_host": (i32, [vp, vp, i64, vp, i32, vp]),
        "trt_dist_wait": (i32, [vp]),
        "trt_dist_finalize": (i32, [vp]),
        "trt_vcf_open": (i32, [C.c_char_p, i32, C.POINTER(vp)]),
        "trt_vcf_close": (None, [vp]),
        "trt_vcf_last_error": (C.c_char_p, [vp]),
        "trt_vcf_header": (i32, [vp, C.POINTER(vp), C.POINTER(i64)]),
        "trt_vcf_n_samples": (i64, [vp]),
        "trt_vcf_set_samples": (i32, [vp, vp, i64]),
        "trt_vcf_seek": (i32, [vp, i64, C.c_int32]),
        "trt_vcf_read_block": (i32, [vp, i64, i64, C.POINTER(vp), C.POINTER(i64)]),
        "trt_vcf_block_free": (None, [vp]),
        "trt_vcf_block_text": (i32, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "trt_vcf_block_parse": (i32, [vp, i32, vp, i32, vp, vp, vp, vp, vp, vp]),
        "trt_vcf_block_parse_packed": (i32, [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]),
        "trt_vcf_block_parse_nibble": (i32, [vp, vp, vp, i32, vp, vp, vp, vp, vp, vp]),
        "trt_vcf_block_field": (i32, [vp, i64, i32, C.c_int32, vp, vp]),
        "trt_vcf_join_samples": (i64, [i64, i32, vp, vp, vp, vp, i64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def device_count() -> int:
    return int(load().trt_device_count())


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dtype):
    """C-contiguous array of the exact dtype (no copy when already so)."""
    return np.ascontiguousarray(a, dtype=dtype)


class Context:
    """One GPU context (``trt_ctx``).  Not thread-safe; one per device per process."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = C.c_void_p()
        rc = self.lib.trt_init(int(device), C.byref(h))
        if rc != TRT_OK:
            msg = self.lib.trt_last_error(None)
            raise TrtError(rc, msg.decode() if msg else "trt_init failed")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.trt_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != TRT_OK:
            msg = self.lib.trt_last_error(self.h)
            raise TrtError(rc, msg.decode() if msg else "error")

    # -- info -----------------------------------------------------------------------------------
    def device_info(self) -> dict:
        d = DevInfo()
        self.check(self.lib.trt_device_info(self.h, C.byref(d)))
        return dict(name=d.name.decode(), cc=(d.cc_major, d.cc_minor), sm_count=d.sm_count,
                    total_mem_bytes=d.total_mem_bytes, free_mem_bytes=d.free_mem_bytes, l2_bytes=d.l2_bytes,
                    abi_version=d.abi_version)

    def launch_count(self) -> int:
        return int(self.lib.trt_launch_count(self.h))

    def last_kernel_ms(self) -> float:
        return float(self.lib.trt_last_kernel_ms(self.h))

    def last_scan_ms(self) -> float:
        return float(self.lib.trt_last_scan_ms(self.h))

    def stopwatch_start(self):
        self.check(self.lib.trt_stopwatch_start(self.h))

    def stopwatch_stop(self) -> float:
        ms = C.c_double()
        self.check(self.lib.trt_stopwatch_stop(self.h, C.byref(ms)))
        return float(ms.value)

    def synchronize(self):
        self.check(self.lib.trt_synchronize(self.h))

    def pinned_empty(self, shape, dtype) -> np.ndarray:
        """numpy array over cudaHostAlloc'd memory (freed when the context closes is NOT automatic:
        keep the returned array's ``_trt_base`` alive and call ``free_pinned``)."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = self.lib.trt_host_alloc(self.h, n)
        if not p:
            self.check(TRT_ENOMEM)
        buf = (C.c_char * max(n, 1)).from_address(p)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        arr.flags.writeable = True
        self._pinned = getattr(self, "_pinned", {})
        self._pinned[arr.ctypes.data] = p
        return arr

    def _result(self, pinned: bool, name: str, shape, dtype) -> np.ndarray:
        """Result buffer of one output.  ``pinned=True`` hands out page-locked arrays cached per (name, shape):
        the device->host copies then run at PCIe speed instead of through the driver's pageable staging, and the
        arrays are REUSED by the next call with the same shapes (copy what must outlive it)."""
        if not pinned:
            return np.empty(shape, dtype)
        shape = tuple(int(x) for x in (shape if isinstance(shape, tuple) else (shape,)))
        key = (name, shape, np.dtype(dtype).str)
        arena = self.__dict__.setdefault("_arena", {})
        if key not in arena:
            for k in [k for k in arena if k[0] == name]:
                self.free_pinned(arena.pop(k))
            arena[key] = self.pinned_empty(shape, dtype)
        return arena[key]

    def free_pinned(self, arr: np.ndarray):
        p = getattr(self, "_pinned", {}).pop(arr.ctypes.data, None)
        if p:
            self.check(self.lib.trt_host_free(self.h, p))

    # -- block ----------------------------------------------------------------------------------
    def block_begin(self, n_loci, n_samples, ploidy, vcftype):
        vt = VCF_TYPES[vcftype] if isinstance(vcftype, str) else int(vcftype)
        self.check(self.lib.trt_block_begin(self.h, int(n_loci), int(n_samples), int(ploidy), vt))
        self.L, self.S, self.P = int(n_loci), int(n_samples), int(ploidy)

    def block_set_gt(self, gt: np.ndarray):
        gt = _c(gt, np.int16)
        assert gt.shape == (self.L, self.S, self.P + 1), (gt.shape, (self.L, self.S, self.P + 1))
        self.check(self.lib.trt_block_set_gt(self.h, _ptr(gt)))

    def block_set_gt_packed(self, gt2: np.ndarray, phase_bits: Optional[np.ndarray] = None):
        """Packed transfer form: uint8 [L][S][2] (allele 0..252, 254 pad, 255 no-call) + optional phase bits
        uint8 [L][ceil(S/8)]; a third of the bytes of ``block_set_gt`` across PCIe, expanded on the device."""
        g = _c(gt2, np.uint8)
        assert g.shape == (self.L, self.S, 2) and self.P == 2, (g.shape, (self.L, self.S, 2), self.P)
        ph = None
        if phase_bits is not None:
            ph = _c(phase_bits, np.uint8)
            assert ph.shape == (self.L, (self.S + 7) // 8)
        self.check(self.lib.trt_block_set_gt_packed(self.h, _ptr(g), _ptr(ph)))

    def block_set_gt_nibble(self, g4: np.ndarray, phase_bits: Optional[np.ndarray] = None):
        """Nibble transfer form: uint8 [L][S], first haplotype in the low nibble (allele 0..13, 14 pad, 15 no-call) +
        optional phase bits; a sixth of the bytes of ``block_set_gt`` across PCIe, expanded on the device."""
        g = _c(g4, np.uint8)
        assert g.shape == (self.L, self.S) and self.P == 2, (g.shape, (self.L, self.S), self.P)
        ph = None
        if phase_bits is not None:
            ph = _c(phase_bits, np.uint8)
            assert ph.shape == (self.L, (self.S + 7) // 8)
        self.check(self.lib.trt_block_set_gt_nibble(self.h, _ptr(g), _ptr(ph)))

    def block_get_gt_nibble(self, locus0, n, out: Optional[np.ndarray] = None, with_phase: bool = False):
        g = np.empty((n, self.S), np.uint8) if out is None else out
        ph = np.empty((n, (self.S + 7) // 8), np.uint8) if with_phase else None
        self.check(self.lib.trt_block_get_gt_nibble(self.h, int(locus0), int(n), _ptr(g), _ptr(ph)))
        return (g, ph) if with_phase else g

    def block_get_gt_packed(self, locus0, n, out: Optional[np.ndarray] = None, with_phase: bool = False):
        g = np.empty((n, self.S, 2), np.uint8) if out is None else out
        ph = np.empty((n, (self.S + 7) // 8), np.uint8) if with_phase else None
        self.check(self.lib.trt_block_get_gt_packed(self.h, int(locus0), int(n), _ptr(g), _ptr(ph)))
        return (g, ph) if with_phase else g

    def block_set_format(self, field_id: int, arr: np.ndarray):
        if arr.dtype.kind == 'i':
            a = _c(arr, np.int32).reshape(self.L, self.S)
            self.check(self.lib.trt_block_set_format_i32(self.h, field_id, _ptr(a)))
        else:
            a = _c(arr, np.float32).reshape(self.L, self.S, -1)
            self.check(self.lib.trt_block_set_format_f32(self.h, field_id, _ptr(a), a.shape[2]))

    def block_set_alleles(self, seqs: bytes, allele_off, locus_off, pos, start, end, period, given_len=None,
                          motifs: Optional[bytes] = None):
        allele_off = _c(allele_off, np.int64)
        locus_off = _c(locus_off, np.int32)
        pos, start, end, period = (_c(x, np.int32) for x in (pos, start, end, period))
        gl = None if given_len is None else _c(given_len, np.float64)
        sbuf = C.create_string_buffer(seqs, len(seqs)) if len(seqs) else None
        mbuf = C.create_string_buffer(motifs, len(motifs)) if motifs else None
        self.nA = int(locus_off[-1]) if len(locus_off) else 0
        self.locus_off = locus_off
        self.check(self.lib.trt_block_set_alleles(
            self.h, C.cast(sbuf, C.c_void_p) if sbuf is not None else None, _ptr(allele_off), _ptr(locus_off),
            _ptr(pos), _ptr(start), _ptr(end), _ptr(period), _ptr(gl),
            C.cast(mbuf, C.c_void_p) if mbuf is not None else None))
        self._period = period

    def harmonize(self) -> dict:
        self.check(self.lib.trt_harmonize(self.h))
        nA, L = self.nA, self.L
        out = dict(allele_len=np.empty(nA, np.float64), trim_off=np.empty(nA, np.int32),
                   trim_len=np.empty(nA, np.int32), len_class=np.empty(nA, np.int32),
                   seq_class=np.empty(nA, np.int32), len_order=np.empty(nA, np.int32),
                   seq_order=np.empty(nA, np.int32), hrun=np.empty(L, np.int32), flags=np.empty(L, np.int32),
                   motif_off=np.empty(L + 1, np.int64))
        mbytes = int(np.sum(np.maximum(self._period, 0)))
        motif = np.empty(max(mbytes, 1), np.uint8)
        ho = HarmonizeOut(**{k: _ptr(v) for k, v in out.items()}, motif=_ptr(motif))
        self.check(self.lib.trt_get_harmonized(self.h, C.byref(ho)))
        out["motif"] = motif[:mbytes].tobytes()
        return out

    def pack_length_genotypes(self) -> np.ndarray:
        self.check(self.lib.trt_pack_length_genotypes(self.h))
        out = np.empty((self.L, self.S, self.P), np.int16)
        self.check(self.lib.trt_get_packed_gt(self.h, _ptr(out)))
        return out

    _STAT_OUT = (("ac", np.int32), ("n_called", np.int64), ("n_called_nonstrict", np.int64), ("n_hom", np.int64),
                 ("n_padded", np.int64), ("thresh", np.float64), ("het", np.float64), ("entropy", np.float64),
                 ("mean", np.float64), ("mode", np.float64), ("var", np.float64), ("hwep", np.float64),
                 ("nalleles", np.int32))

    def locus_stats(self, use_length: bool, group_masks: Optional[np.ndarray] = None,
                    nalleles_thresh: float = 0.01, want=None, pinned: bool = False) -> dict:
        G = 1 if group_masks is None else int(group_masks.shape[0])
        gm = None if group_masks is None else _c(group_masks, np.uint8).reshape(G, self.S)
        L, nA = self.L, self.nA
        out = {k: self._result(pinned, "stat_" + k, (G, nA if k == "ac" else L), dt)
               for k, dt in self._STAT_OUT if want is None or k in want}
        so = LocusStatsOut(**{k: _ptr(v) for k, v in out.items()})
        self.check(self.lib.trt_locus_stats(self.h, 1 if use_length else 0, _ptr(gm), G, float(nalleles_thresh),
                                            C.byref(so)))
        return out

    def genotype_counts(self, locus: int, n_alleles: int, mask: Optional[np.ndarray] = None) -> np.ndarray:
        """dense (A+2)^P table of sorted-index genotype counts of one locus (digit = allele + 2)."""
        n = (n_alleles + 2) ** self.P
        table = np.zeros(n, np.int64)
        m = None if mask is None else _c(mask, np.uint8)
        self.check(self.lib.trt_genotype_counts(self.h, int(locus), _ptr(m), _ptr(table), n))
        return table.reshape((n_alleles + 2,) * self.P)

    def call_filters(self, specs, dp_field, filter_counts, numcalls, totaldp, want_mask=True, want_trigger=False,
                     want_gt=True) -> dict:
        """ApplyCallFilters on the current block.  ``specs``: list of (kind, field_id, threshold) in filter
        order.  The per-sample accumulators (int64 [n_specs, S], int64 [S], float64 [S]) are updated in place."""
        n = len(specs)
        arr = (CallFilterSpec * max(n, 1))()
        for i, (kind, fld, thr) in enumerate(specs):
            arr[i].kind, arr[i].field_id, arr[i].threshold = int(kind), int(fld), float(thr)
        L, S, P = self.L, self.S, self.P
        assert filter_counts.dtype == np.int64 and filter_counts.flags.c_contiguous and filter_counts.shape == (n, S)
        assert numcalls.dtype == np.int64 and totaldp.dtype == np.float64
        res = {}
        if want_mask:
            res["call_mask"] = np.empty((L, S), np.uint32)
        if want_trigger and n:
            res["trigger_values"] = np.empty((n, L, S), np.float64)
        if want_gt:
            res["gt_masked"] = np.empty((L, S, P + 1), np.int16)
        neg = np.full(1, -1, np.int32)
        out = CallFilterOut(call_mask=_ptr(res.get("call_mask")), trigger_values=_ptr(res.get("trigger_values")),
                            gt_masked=_ptr(res.get("gt_masked")), filter_counts=_ptr(filter_counts) if n else None,
                            numcalls=_ptr(numcalls), totaldp=_ptr(totaldp), negative_dp_locus=_ptr(neg))
        self.check(self.lib.trt_call_filters(self.h, arr, n, int(dp_field), C.byref(out)))
        res["negative_dp_locus"] = int(neg[0])
        return res

    def locus_filters(self, specs, use_length: bool, pinned: bool = False) -> dict:
        """ApplyLocusFilters + INFO recompute on the current block.  ``specs``: list of (kind, threshold)."""
        n = len(specs)
        arr = (LocusFilterSpec * max(n, 1))()
        for i, (kind, thr) in enumerate(specs):
            arr[i].kind, arr[i].threshold = int(kind), float(0.0 if thr is None else thr)
        L, nA = self.L, self.nA
        r = lambda k, n_, dt: self._result(pinned, "lf_" + k, (n_,), dt)
        res = dict(flags=r("flags", L, np.uint32), n_called=r("n_called", L, np.int64), het=r("het", L, np.float64),
                   hwep=r("hwep", L, np.float64), ac=r("ac", nA, np.int32), hrun=r("hrun", L, np.int32))
        out = LocusFilterOut(**{k: _ptr(v) for k, v in res.items()})
        self.check(self.lib.trt_locus_filters(self.h, arr, n, 1 if use_length else 0, C.byref(out)))
        return res

    def assoc_set_design(self, covars: np.ndarray, outcome: np.ndarray, sample_index: np.ndarray):
        """Standardised design (associaTR.py:190-194): covars float64 [n, K] with column 0 reserved for the
        genotype and column 1 = 1; outcome float64 [n]; sample_index int32 [n] = VCF sample of each row."""
        cv = _c(covars, np.float64)
        oc = _c(outcome, np.float64)
        si = _c(sample_index, np.int32)
        assert cv.ndim == 2 and cv.shape[0] == oc.shape[0] == si.shape[0]
        self.check(self.lib.trt_assoc_set_design(self.h, _ptr(cv), _ptr(oc), _ptr(si), cv.shape[0], cv.shape[1]))

    def assoc_ols(self, non_major_cutoff: float, pinned: bool = False, want=None) -> dict:
        """``want=()``: leave the rows in device memory (multi-GPU runs gather them with trt_dist_gather_region)."""
        L, nA = self.L, self.nA
        r = lambda k, n_, dt: self._result(pinned, "assoc_" + k, (n_,), dt)
        spec = (("filter_code", L, np.int32), ("n_tested", L, np.int64), ("p", L, np.float64), ("coef", L, np.float64),
                ("se", L, np.float64), ("r2", L, np.float64), ("std_g", L, np.float64), ("ac_len", nA, np.int32))
        res = {k: r(k, n_, dt) for k, n_, dt in spec if want is None or k in want}
        out = AssocOut(**{k: _ptr(v) for k, v in res.items()})
        self.check(self.lib.trt_assoc_ols(self.h, float(non_major_cutoff), C.byref(out)))
        return res

    def block_set_ap(self, ap1: np.ndarray, ap2: np.ndarray, has_ap: Optional[np.ndarray] = None):
        """Beagle FORMAT AP1 / AP2 of the block: float32, the records' [S][A-1] arrays concatenated."""
        a1, a2 = _c(ap1, np.float32).reshape(-1), _c(ap2, np.float32).reshape(-1)
        n = self.S * (self.nA - self.L)
        assert a1.size == n and a2.size == n, (a1.size, a2.size, n)
        h = None if has_ap is None else _c(has_ap, np.uint8)
        self.check(self.lib.trt_block_set_ap(self.h, _ptr(a1), _ptr(a2), _ptr(h)))

    def dosages(self, dosage_type) -> tuple:
        """TRRecord.GetDosages for every locus: (float32 [L, S], int32 [L] record-level validation codes DE_*)."""
        t = DOSAGE_TYPES[dosage_type] if isinstance(dosage_type, str) else int(dosage_type)
        out = np.empty((self.L, self.S), np.float32)
        err = np.zeros(self.L, np.int32)
        self.check(self.lib.trt_dosages(self.h, t, _ptr(out), _ptr(err)))
        return out, err

    def assoc_dosage_ols(self, cls: np.ndarray, len_round: np.ndarray, len_around: np.ndarray) -> dict:
        """associaTR --beagle-dosages on the current block (see include/trtools_b200.h trt_assoc_dosage_ols)."""
        L, nA = self.L, self.nA
        c, lr, la = _c(cls, np.int32), _c(len_round, np.float64), _c(len_around, np.float64)
        assert c.size == nA and lr.size == nA and la.size == nA
        res = dict(n_tested=np.empty(L, np.int64), p=np.empty(L), coef=np.empty(L), se=np.empty(L), r2=np.empty(L),
                   std_g=np.empty(L), ncovars_code=np.empty(L, np.int32), class_stats=np.empty((nA, 4)),
                   length_stats=np.empty((L, 5)))
        out = AssocDosageOut(**{k: _ptr(v) for k, v in res.items()})
        self.check(self.lib.trt_assoc_dosage_ols(self.h, _ptr(c), _ptr(lr), _ptr(la), C.byref(out)))
        return res

    def qc_reduce(self, sample_calls: np.ndarray, sample_quality: Optional[np.ndarray], sample_mask: Optional[np.ndarray],
                  quality_field: int = -1, ignore_no_call: bool = False, rec_ploidy: Optional[np.ndarray] = None) -> dict:
        """qcSTR's per-block reductions (trt_qc_reduce).  ``sample_calls`` int64 [S] and ``sample_quality`` float64 [S]
        are accumulated in place over the whole sample axis (entries outside ``sample_mask`` stay untouched)."""
        assert sample_calls.dtype == np.int64 and sample_calls.shape == (self.S,)
        m = None if sample_mask is None else _c(sample_mask, np.uint8)
        res = dict(locus_calls=np.empty(self.L, np.int64), locus_quality=np.empty(self.L, np.float64))
        out = QcOut(sample_calls=_ptr(sample_calls), locus_calls=_ptr(res["locus_calls"]),
                    sample_quality=_ptr(sample_quality) if quality_field >= 0 else None,
                    locus_quality=_ptr(res["locus_quality"]) if quality_field >= 0 else None)
        rp = None if rec_ploidy is None else _c(rec_ploidy, np.int32)
        assert rp is None or rp.shape == (self.L,)
        self.check(self.lib.trt_qc_reduce(self.h, _ptr(m), _ptr(rp), int(quality_field), 1 if ignore_no_call else 0, C.byref(out)))
        return res

    def compare(self, gt2: np.ndarray, idx1, idx2, locus_off2, seq_id2, len2, reflen, ignore_phasing: bool,
                sample_numcalls: np.ndarray, sample_conc_seq: np.ndarray, sample_conc_len: np.ndarray) -> dict:
        """compareSTR.UpdateComparisonResults for the current block against a second call set (trt_compare)."""
        g2 = _c(gt2, np.int16)
        assert g2.ndim == 3 and g2.shape[0] == self.L and g2.shape[2] == self.P + 1, (g2.shape, self.L, self.P)
        i1, i2 = _c(idx1, np.int32), _c(idx2, np.int32)
        lo2, sid, l2, rl = _c(locus_off2, np.int32), _c(seq_id2, np.int32), _c(len2, np.float64), _c(reflen, np.float64)
        n = len(i1)
        for a in (sample_numcalls, sample_conc_seq, sample_conc_len):
            assert a.dtype == np.int64 and a.shape == (n,)
        res = dict(numcalls=np.empty(self.L, np.int64), conc_seq=np.empty(self.L, np.int64), conc_len=np.empty(self.L, np.int64),
                   len_sums=np.empty((self.L, 5)), status=np.empty(self.L, np.int32))
        cin = CompareIn(gt2=_ptr(g2), S2=g2.shape[1], idx1=_ptr(i1), idx2=_ptr(i2), n_shared=n, locus_off2=_ptr(lo2),
                        seq_id2=_ptr(sid), len2=_ptr(l2), reflen=_ptr(rl), ignore_phasing=1 if ignore_phasing else 0)
        out = CompareOut(sample_numcalls=_ptr(sample_numcalls), sample_conc_seq=_ptr(sample_conc_seq),
                         sample_conc_len=_ptr(sample_conc_len), **{k: _ptr(v) for k, v in res.items()})
        self.check(self.lib.trt_compare(self.h, C.byref(cin), C.byref(out)))
        return res

    def synth_fill(self, seed, locus_offset, cum_freq, miss_thresh, half_thresh, with_format=True):
        cf = _c(cum_freq, np.uint32)
        self.check(self.lib.trt_synth_fill(self.h, int(seed), int(locus_offset), self.L, self.S, _ptr(cf),
                                           int(miss_thresh), int(half_thresh), int(with_format) if not isinstance(with_format, bool) else (15 if with_format else 0)))

    def block_get_gt(self, locus0, n) -> np.ndarray:
        out = np.empty((n, self.S, self.P + 1), np.int16)
        self.check(self.lib.trt_block_get_gt(self.h, int(locus0), int(n), _ptr(out)))
        return out

    def block_get_format(self, field_id, locus0, n, dtype, ncol=1) -> np.ndarray:
        out = np.empty((n, self.S, ncol) if ncol > 1 else (n, self.S), dtype)
        self.check(self.lib.trt_block_get_format(self.h, int(field_id), int(locus0), int(n), _ptr(out)))
        return out


_default_ctx = {}


def default_context(device: Optional[int] = None) -> Context:
    """Process-wide context for ``device`` (default: LOCAL_RANK or 0)."""
    if device is None:
        device = int(os.environ.get("TRTOOLS_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]

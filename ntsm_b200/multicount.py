"""Python mirrors of the reference's MultiCount (src/MultiCount.hpp) and VCFConvert (src/VCFConvert.hpp)
over the C ABI -- the multi-sample matrix path.

    mc = MultiCount(sites_path, sample_ids, k=19, dupes=False)     # MultiCount(sampleIDs)        :43
    mc.insertCount(sample_index, hash_value, multi)                # insertCount                  :52
    mc.insertWindows(windows, genotypes, multi)                    # VCFConvert::count inner loops, batched
    text = mc.printCountsMax(index)                                # printCountsMax               :93
    mc.printNormMatrix(matrix_path, center_path)                   # printNormMatrix              :148

    vc = VCFConvert(sites_path, ref_fasta, k=19, multi=20, window=31)   # VCFConvert()            :42
    vc.count(vcf_path)                                             # count                        :62
    vc.outputMatrix(prefix); vc.outputCounts(dir)                  # outputMatrix / outputCounts  :186 / :173

Everything that touches the matrix runs in libntsm_b200.so on the GPU.  KeyError is what the reference's
uncaught std::out_of_range / failed assert become here.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import NtsmError, check
from .fingerprint import FingerPrint, SiteSet

UNDEF = float(np.finfo(np.float64).max)      # MultiCount::UNDEF (:41)


def _raise(rc, ctx):
    if rc == -134:
        raise KeyError((_lib.lib().ntsm_last_error(ctx) or b"Couldn't find key.").decode(errors="replace"))
    check(rc, ctx)


class MultiCount:
    def __init__(self, sites, sample_ids, k=19, dupes=False, device=0, _fp=None, _handle=None):
        L = _lib.lib()
        self._fp = _fp or FingerPrint(sites if isinstance(sites, SiteSet) else SiteSet(sites, k, dupes), k=k, device=device, batch_bases=4096, n_buffers=2)
        self.sites = self._fp.sites
        self.sample_ids = list(sample_ids)
        self._owned = _handle is None
        if _handle is None:
            self._h = C.c_void_p()
            check(L.ntsm_multi_create(C.byref(self._h), self._fp._ctx, len(self.sample_ids)), self._fp._ctx)
        else:
            self._h = _handle

    def close(self):
        if getattr(self, "_h", None) and self._owned:
            _lib.lib().ntsm_multi_destroy(self._h)
        self._h = None

    __del__ = close

    @property
    def launches(self):
        return self._fp.launches

    def insertCount(self, sample_index, hash_value, multi=1):
        check(_lib.lib().ntsm_multi_insert_count(self._h, sample_index, hash_value, multi), self._fp._ctx)

    def insertWindows(self, windows, genotypes, multi=20):
        """windows: [(ref_window_bytes, alt_window_bytes), ...] one pair per SNP line; genotypes: array
        [n_lines][n_samples] of 0 (hom ref / unknown), 1 (het), 2 (hom alt)."""
        n = len(windows)
        stride = max([16] + [len(w) for pair in windows for w in pair])
        stride = (stride + 15) & ~15
        buf = np.zeros((n, 2, stride), np.uint8)
        lens = np.zeros((n, 2), np.uint16)
        for i, pair in enumerate(windows):
            for a in range(2):
                w = np.frombuffer(pair[a], np.uint8)
                buf[i, a, :len(w)] = w
                lens[i, a] = len(w)
        g = np.ascontiguousarray(genotypes, np.uint8).reshape(n, len(self.sample_ids)) if n else np.zeros((0, 0), np.uint8)
        _raise(_lib.lib().ntsm_multi_insert_windows(self._h, buf.ctypes.data, stride, lens.ctypes.data, g.ctypes.data, n, multi), self._fp._ctx)

    @property
    def warnings_text(self):
        L = _lib.lib()
        n = L.ntsm_multi_warnings_text(self._h, None, 0)
        b = C.create_string_buffer(max(1, n))
        L.ntsm_multi_warnings_text(self._h, b, n)
        return b.raw[:n]

    def matrix(self):
        out = np.zeros((len(self.sample_ids), self.sites.n_kmers), np.uint8)
        if out.size:
            check(_lib.lib().ntsm_multi_get_matrix(self._h, out.ctypes.data), self._fp._ctx)
        return out

    def countsMax(self, index):
        S = self.sites.n_sites
        a = [np.zeros(S, np.uint32) for _ in range(4)]
        check(_lib.lib().ntsm_multi_counts_max(self._h, index, *[x.ctypes.data for x in a]), self._fp._ctx)
        return a

    def printCountsMax(self, index):
        L = _lib.lib()
        n = L.ntsm_multi_format_counts(self._h, self.sites._h, index, None, 0)
        _raise(n, self._fp._ctx)
        b = C.create_string_buffer(max(1, n))
        L.ntsm_multi_format_counts(self._h, self.sites._h, index, b, n)
        return b.raw[:n].decode()

    def normMatrix(self):
        """(values[n_sites][n_samples] with UNDEF for missing, sums[n_sites])"""
        S, N = self.sites.n_sites, len(self.sample_ids)
        v = np.zeros((S, N), np.float64)
        s = np.zeros(S, np.float64)
        check(_lib.lib().ntsm_multi_norm_matrix(self._h, v.ctypes.data if v.size else None, s.ctypes.data if S else None), self._fp._ctx)
        return v, s

    def printNormMatrix(self, matrix_path, center_path, threads=1):
        ids = (C.c_char_p * (len(self.sample_ids) + 1))(*[s.encode() if isinstance(s, str) else s for s in self.sample_ids], None)
        _raise(_lib.lib().ntsm_multi_write_norm_matrix(self._h, self.sites._h, ids, os.fsencode(matrix_path), os.fsencode(center_path), threads), self._fp._ctx)


class VCFConvert:
    def __init__(self, sites, ref, k=19, dupes=False, multi=20, window=31, threads=1, device=0, verbose=0):
        self._fp = FingerPrint(sites if isinstance(sites, SiteSet) else SiteSet(sites, k, dupes), k=k, device=device, batch_bases=4096, n_buffers=2)
        self.ref, self.multi, self.window, self.threads, self.verbose = ref, multi, window, threads, verbose
        self._h = None
        self.counts = None          # m_counts, after count()

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().ntsm_vcf_destroy(self._h)
            self._h = None

    __del__ = close

    def count(self, filename):
        L = _lib.lib()
        self.close()
        h = C.c_void_p()
        rc = L.ntsm_vcf_convert(C.byref(h), self._fp._ctx, self._fp.sites._h, os.fsencode(self.ref), os.fsencode(filename), self.multi,
                                self.window, self.threads, self.verbose)
        if rc == -5:
            raise FileNotFoundError((L.ntsm_last_error(self._fp._ctx) or b"").decode())
        _raise(rc, self._fp._ctx)
        self._h = h
        self.sample_ids = [L.ntsm_vcf_sample_id(h, i).decode() for i in range(L.ntsm_vcf_n_samples(h))]
        self.lines_counted = L.ntsm_vcf_lines_counted(h)
        self.counts = MultiCount(self._fp.sites, self.sample_ids, _fp=self._fp, _handle=C.c_void_p(L.ntsm_vcf_multi(h)))

    def outputMatrix(self, prefix):
        _raise(_lib.lib().ntsm_vcf_output_matrix(self._h, os.fsencode(prefix)), self._fp._ctx)

    def outputCounts(self, directory=None):
        _raise(_lib.lib().ntsm_vcf_output_counts(self._h, os.fsencode(directory) if directory else None), self._fp._ctx)


def parse_vcf(ref, vcf, window=31, threads=1, verbose=0):
    """The host half of VCFConvert::count on its own (ntsm_vcf_parse; no GPU needed): returns
    (rc, sample_ids, [(ref_window, alt_window), ...], genotypes[n_lines][n_samples]).  rc is 0, or the code
    ntsm_vcf_convert would return (-134 where the reference dies); the lines in front of a fatal one are still returned."""
    L = _lib.lib()
    h = C.c_void_p()
    rc = L.ntsm_vcf_parse(C.byref(h), os.fsencode(ref), os.fsencode(vcf), window, threads, verbose)
    if not h:
        return rc, [], [], np.zeros((0, 0), np.uint8)
    try:
        S, n, ws = L.ntsm_vcf_lines_n_samples(h), L.ntsm_vcf_lines_count(h), L.ntsm_vcf_lines_wstride(h)
        ids = [L.ntsm_vcf_lines_sample_id(h, i).decode() for i in range(S)]
        wins = []
        if n:
            raw = np.ctypeslib.as_array(C.cast(L.ntsm_vcf_lines_windows(h), C.POINTER(C.c_uint8)), (2 * n, ws))
            lens = np.ctypeslib.as_array(C.cast(L.ntsm_vcf_lines_lens(h), C.POINTER(C.c_uint16)), (2 * n,))
            wins = [(raw[2 * i, :lens[2 * i]].tobytes(), raw[2 * i + 1, :lens[2 * i + 1]].tobytes()) for i in range(n)]
        g = np.zeros((n, S), np.uint8)
        if n and S:
            g = np.ctypeslib.as_array(C.cast(L.ntsm_vcf_lines_genotypes(h), C.POINTER(C.c_uint8)), (n, S)).copy()
        return rc, ids, wins, g
    finally:
        L.ntsm_vcf_lines_free(h)

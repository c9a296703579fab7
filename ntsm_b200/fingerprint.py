"""Python mirror of the reference's FingerPrint object (src/FingerPrint.hpp) over the C ABI.

    fp = FingerPrint(sites_path, k=19, dupes=False, cov_thresh=0)   # FingerPrint()        :35
    fp.computeCounts(files, threads=1)                               # computeCounts        :46
    fp.insertCount(seq)                                              # insertCount          :89
    text = fp.counts_text()                  # printOptionalHeader + printCountsMax         :261-311
    info = fp.printInfoSummary()                                     # printInfoSummary     :313

Names, argument meaning and error behaviour follow the reference; everything that counts runs in
libntsm_b200.so on the GPU.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import Cfg, NtsmError, check


class SiteSet:
    """FingerPrint::initCountsHash (src/FingerPrint.hpp:490-564): the site table on the host."""

    def __init__(self, path, k=19, dupes=False):
        L = _lib.lib()
        self._h = C.c_void_p()
        rc = L.ntsm_sites_load(C.byref(self._h), os.fsencode(path), k, int(dupes))
        if rc == -5:
            raise FileNotFoundError("file %s cannot be opened" % path)      # FingerPrint.hpp:493-499
        check(rc)
        self.k = k
        self.n_sites = L.ntsm_sites_n_sites(self._h)
        self.n_kmers = L.ntsm_sites_n_kmers(self._h)
        self.table_size = L.ntsm_sites_table_size(self._h)

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().ntsm_sites_free(self._h)
            self._h = None

    __del__ = close

    @property
    def hashes(self):
        return np.ctypeslib.as_array(_lib.lib().ntsm_sites_hashes(self._h), (self.n_kmers,)).copy() if self.n_kmers else np.zeros(0, np.uint64)

    @property
    def allele_off(self):
        return np.ctypeslib.as_array(_lib.lib().ntsm_sites_allele_off(self._h), (2 * self.n_sites + 1,)).copy()

    @property
    def erased(self):
        return np.ctypeslib.as_array(_lib.lib().ntsm_sites_erased(self._h), (self.n_kmers,)).copy() if self.n_kmers else np.zeros(0, np.uint8)

    @property
    def names(self):
        L = _lib.lib()
        return [L.ntsm_sites_name(self._h, i).decode() for i in range(self.n_sites)]

    @property
    def warnings(self):
        L = _lib.lib()
        return [L.ntsm_sites_warning(self._h, i).decode() for i in range(L.ntsm_sites_n_warnings(self._h))]

    def printable(self):
        return _lib.lib().ntsm_sites_printable(self._h) == 0

    def max_counts(self, cov):
        return _lib.lib().ntsm_sites_max_counts(self._h, float(cov))


def pack_reads(reads):
    """Pack a list of byte strings with the library's packer -> (bases2 u32[], nmask u32[], n_pos, read_off)."""
    L = _lib.lib()
    buf = b"".join(reads)
    off = np.zeros(len(reads) + 1, np.uint64)
    np.cumsum([len(r) for r in reads], out=off[1:])
    n_pos_max = sum((len(r) + 8) & ~7 for r in reads)        # bases + separator, rounded up to 8 positions per read
    padded = L.ntsm_padded_positions(n_pos_max)
    bases = np.zeros(padded // 16, np.uint32)
    mask = np.zeros(padded // 32, np.uint32)
    roff = np.zeros(len(reads) + 1, np.uint64)
    cbuf = C.create_string_buffer(buf, len(buf) + 1)
    n_pos = L.ntsm_pack_reads(C.cast(cbuf, C.c_void_p), off.ctypes.data, len(reads), bases.ctypes.data, mask.ctypes.data, roff.ctypes.data)
    return bases, mask, int(n_pos), roff


class FingerPrint:
    """One GPU context with a loaded site table (the FingerPrint object of the reference)."""

    def __init__(self, sites, k=19, dupes=False, cov_thresh=0.0, device=0, batch_bases=0, n_buffers=0, options=None):
        """options: {name: int} for ntsm_ctx_set_option (measurement / test knobs: "kernel", "pair_fold",
        "filter_bits", "launch_shape", "l2_persist", "device_pack"); every setting gives the same counts."""
        L = _lib.lib()
        self.sites = sites if isinstance(sites, SiteSet) else SiteSet(sites, k, dupes)
        self.k = self.sites.k
        cfg = Cfg(k=self.k, device=device, n_buffers=n_buffers, reserved=0, batch_bases=batch_bases,
                  max_counts=self.sites.max_counts(cov_thresh))
        self.max_counts = cfg.max_counts
        self._ctx = C.c_void_p()
        check(L.ntsm_ctx_create(C.byref(self._ctx), C.byref(cfg)))
        self.options = dict(options or {})
        for name, value in self.options.items():
            check(L.ntsm_ctx_set_option(self._ctx, name.encode(), int(value)), self._ctx)
        check(L.ntsm_load_siteset(self._ctx, self.sites._h), self._ctx)
        self.early_term = False
        self._rows = None

    def close(self):
        if getattr(self, "_ctx", None):
            _lib.lib().ntsm_ctx_destroy(self._ctx)
            self._ctx = None

    __del__ = close

    # -- counting -------------------------------------------------------------------
    def insertCount(self, seq: bytes):
        """FingerPrint::insertCount(seq, len) -- src/FingerPrint.hpp:89."""
        check(_lib.lib().ntsm_insert_count(self._ctx, seq, len(seq)), self._ctx)

    def insertReads(self, buf, off, threads=1):
        """Bulk insertCount: read r = buf[off[r]:off[r+1]] (ntsm_insert_reads).  buf: bytes / uint8 array, off: uint64 array."""
        b = np.frombuffer(buf, np.uint8) if isinstance(buf, (bytes, bytearray)) else buf
        off = np.ascontiguousarray(off, np.uint64)
        ctxs = (C.c_void_p * 1)(self._ctx)
        rc = _lib.lib().ntsm_insert_reads(ctxs, 1, b.ctypes.data if len(b) else None, off.ctypes.data, len(off) - 1, threads)
        check(rc, self._ctx if rc != -1 else None)

    def insertReadsFixed(self, ptr, read_len, stride, n_reads, threads=1):
        """Bulk insertCount over a dense host matrix of n_reads x read_len ASCII bytes at address `ptr`."""
        ctxs = (C.c_void_p * 1)(self._ctx)
        check(_lib.lib().ntsm_insert_reads_fixed(ctxs, 1, ptr, read_len, stride, n_reads, threads), self._ctx)

    def computeCounts(self, filenames, threads=1, verbose=0):
        """FingerPrint::computeCounts -- src/FingerPrint.hpp:46; `threads` is opt::threads."""
        L = _lib.lib()
        arr = (C.c_char_p * len(filenames))(*[os.fsencode(f) for f in filenames])
        ctxs = (C.c_void_p * 1)(self._ctx)
        early = C.c_int(0)
        rc = L.ntsm_count_files(ctxs, 1, arr, len(filenames), threads, verbose, C.byref(early))
        if rc == -5:
            raise FileNotFoundError(L.ntsm_last_error(None).decode())
        check(rc)
        self.early_term = bool(early.value)

    def count_packed_device(self, d_bases_ptr, d_mask_ptr, n_pos, n_bases, stream=None):
        """Count a packed stream already resident in device memory (see ntsm_count_packed_device)."""
        check(_lib.lib().ntsm_count_packed_device(self._ctx, d_bases_ptr, d_mask_ptr, n_pos, n_bases, stream), self._ctx)

    def count_packed_host(self, h_bases_ptr, h_mask_ptr, n_pos, n_bases):
        """Count a packed stream in (pinned) host memory: H2D slices overlap the kernel (ntsm_count_packed_host)."""
        check(_lib.lib().ntsm_count_packed_host(self._ctx, h_bases_ptr, h_mask_ptr, n_pos, n_bases), self._ctx)

    def set_stream(self, stream):
        check(_lib.lib().ntsm_set_stream(self._ctx, stream), self._ctx)

    def reset_async(self):
        check(_lib.lib().ntsm_reset_counts_async(self._ctx), self._ctx)
        self._rows = None

    def reduce_async(self):
        check(_lib.lib().ntsm_reduce_async(self._ctx), self._ctx)

    def flush(self):
        check(_lib.lib().ntsm_flush(self._ctx), self._ctx)

    def sync(self):
        check(_lib.lib().ntsm_sync(self._ctx), self._ctx)

    def reset(self):
        check(_lib.lib().ntsm_reset_counts(self._ctx), self._ctx)
        self._rows = None
        self.early_term = False

    def poll_totals(self):
        a, b, c, d = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_int()
        check(_lib.lib().ntsm_poll_totals(self._ctx, C.byref(a), C.byref(b), C.byref(c), C.byref(d)), self._ctx)
        return a.value, b.value, c.value, bool(d.value)

    def set_option(self, name, value):
        check(_lib.lib().ntsm_ctx_set_option(self._ctx, name.encode(), int(value)), self._ctx)

    # -- multi GPU --------------------------------------------------------------------
    @staticmethod
    def group_finalize(fps):
        """Several FingerPrints of this process (one per GPU) combined without NCCL: fps[0]'s GPU sums the
        others' counts out of peer memory inside the per-site reduce kernel (ntsm_group_finalize).
        Returns the rows like finalize(); they are also what fps[0].counts_text() then prints."""
        f0 = fps[0]
        S = f0.sites.n_sites
        a = [np.zeros(max(S, 1), np.uint32) for _ in range(4)]
        t = np.zeros(3, np.uint64)
        ctxs = (C.c_void_p * len(fps))(*[f._ctx for f in fps])
        check(_lib.lib().ntsm_group_finalize(ctxs, len(fps), a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data,
                                             t.ctypes.data), f0._ctx)
        f0._rows = [x[:S] for x in a] + [t]
        return f0._rows

    @staticmethod
    def nccl_unique_id():
        buf = C.create_string_buffer(128)
        check(_lib.lib().ntsm_nccl_unique_id(buf))
        return buf.raw

    def comm_init(self, uid: bytes, rank, n_ranks):
        check(_lib.lib().ntsm_comm_init(self._ctx, uid, rank, n_ranks), self._ctx)

    def allreduce(self):
        check(_lib.lib().ntsm_allreduce(self._ctx), self._ctx)

    # -- results ----------------------------------------------------------------------
    def finalize(self):
        """Drain, combine, per-site reduce.  Returns (max_ref, max_var, sum_ref, sum_var, totals[3])."""
        S = self.sites.n_sites
        a = [np.zeros(max(S, 1), np.uint32) for _ in range(4)]
        t = np.zeros(3, np.uint64)
        check(_lib.lib().ntsm_finalize(self._ctx, a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, t.ctypes.data), self._ctx)
        self._rows = [x[:S] for x in a] + [t]
        return self._rows

    def kmer_counts(self):
        """m_counts values in dense k-mer order (site list order)."""
        out = np.zeros(max(self.sites.n_kmers, 1), np.uint32)
        check(_lib.lib().ntsm_get_counts(self._ctx, out.ctypes.data), self._ctx)
        return out[:self.sites.n_kmers]

    def counts_text(self):
        """printOptionalHeader() + printCountsMax() -- the counts file, byte for byte (:261-311)."""
        L = _lib.lib()
        mr, mv, sr, sv, t = self._rows or self.finalize()
        args = (self.sites._h, mr.ctypes.data, mv.ctypes.data, sr.ctypes.data, sv.ctypes.data, int(t[0]))
        n = L.ntsm_format_counts(*args, None, 0)
        if n == -134:
            raise KeyError("Couldn't find key.")      # std::out_of_range from m_counts.at(), :282
        buf = C.create_string_buffer(n + 1)
        L.ntsm_format_counts(*args, buf, n)
        return buf.raw[:n].decode()

    def sites_covered(self):
        mr, mv, _, _, _ = self._rows or self.finalize()
        return _lib.lib().ntsm_sites_covered(mr.ctypes.data, mv.ctypes.data, self.sites.n_sites)

    def printInfoSummary(self):
        """printInfoSummary() text -- src/FingerPrint.hpp:313-333."""
        L = _lib.lib()
        _, _, _, _, t = self._rows or self.finalize()
        buf = C.create_string_buffer(2048)
        n = L.ntsm_format_summary(self.sites._h, t.ctypes.data, self.sites_covered(), buf, 2048)
        return buf.raw[:n].decode()

    @property
    def launches(self):
        return _lib.lib().ntsm_ctx_launches(self._ctx)

    @property
    def kernel_name(self):
        """Name of the count kernel this context launches (as ncu lists it)."""
        return _lib.lib().ntsm_ctx_kernel_name(self._ctx).decode()

    @property
    def filter_bits(self):
        return _lib.lib().ntsm_ctx_filter_bits(self._ctx)

    @property
    def pcie_bytes(self):
        """(host->device, device->host) bytes this context's data path has copied so far."""
        a, b = C.c_uint64(), C.c_uint64()
        _lib.lib().ntsm_ctx_pcie_bytes(self._ctx, C.byref(a), C.byref(b))
        return a.value, b.value

    @property
    def l2_window(self):
        """0 = launches carry no L2 access-policy window, else the window's hit ratio in percent."""
        return _lib.lib().ntsm_ctx_l2_window(self._ctx)

    @property
    def probe_bytes(self):
        return _lib.lib().ntsm_ctx_probe_bytes(self._ctx)

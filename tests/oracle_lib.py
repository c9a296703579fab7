"""ctypes view of oracle/_build/libntsm_oracle.so -- the CPU checker (test infrastructure)."""
import ctypes as C
import io
import os
import tempfile

import numpy as np

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


class Oracle:
    def __init__(self, path):
        L = self.L = C.CDLL(path)
        L.ntsm_oracle_nt4.restype = C.c_int
        L.ntsm_oracle_nt4.argtypes = [C.c_ubyte]
        L.ntsm_oracle_hash64.restype = C.c_uint64
        L.ntsm_oracle_hash64.argtypes = [C.c_uint64, C.c_uint]
        L.ntsm_oracle_iter.restype = C.c_size_t
        L.ntsm_oracle_iter.argtypes = [C.c_char_p, C.c_uint64, C.c_uint, u64p, u64p, u64p, u64p, C.c_size_t]
        L.ntsm_oracle_fp_create.restype = C.c_void_p
        L.ntsm_oracle_fp_create.argtypes = [C.c_char_p, C.c_uint, C.c_int, C.c_double, C.c_void_p]
        L.ntsm_oracle_fp_destroy.argtypes = [C.c_void_p]
        L.ntsm_oracle_fp_insert.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.ntsm_oracle_fp_insert_many.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
        L.ntsm_oracle_fp_count_file.restype = C.c_int
        L.ntsm_oracle_fp_count_file.argtypes = [C.c_void_p, C.c_char_p]
        L.ntsm_oracle_fp_summary.restype = C.c_int
        L.ntsm_oracle_fp_summary.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
        for f in ("total_kmers", "total_counts", "total_bases", "max_counts", "table_size"):
            getattr(L, "ntsm_oracle_fp_" + f).restype = C.c_uint64
            getattr(L, "ntsm_oracle_fp_" + f).argtypes = [C.c_void_p]
        for f in ("n_sites", "n_listed"):
            getattr(L, "ntsm_oracle_fp_" + f).restype = C.c_uint32
            getattr(L, "ntsm_oracle_fp_" + f).argtypes = [C.c_void_p]
        L.ntsm_oracle_fp_early_term.restype = C.c_int
        L.ntsm_oracle_fp_early_term.argtypes = [C.c_void_p]
        L.ntsm_oracle_fp_rows.restype = C.c_int
        L.ntsm_oracle_fp_rows.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        L.ntsm_oracle_fp_lists.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ntsm_oracle_fp_site_name.restype = C.c_char_p
        L.ntsm_oracle_fp_site_name.argtypes = [C.c_void_p, C.c_uint32]
        L.ntsm_oracle_reader_open.restype = C.c_void_p
        L.ntsm_oracle_reader_open.argtypes = [C.c_char_p]
        L.ntsm_oracle_reader_next.restype = C.c_long
        L.ntsm_oracle_reader_next.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p)]
        L.ntsm_oracle_reader_close.argtypes = [C.c_void_p]
        # multi-sample matrix path (MultiCount / VCFConvert)
        L.ntsm_oracle_vcf_run.restype = C.c_int
        L.ntsm_oracle_vcf_run.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_uint, C.c_int, C.c_uint, C.c_uint, C.c_char_p]
        L.ntsm_oracle_mc_create.restype = C.c_void_p
        L.ntsm_oracle_mc_create.argtypes = [C.c_char_p, C.c_uint, C.c_int, C.c_uint32, C.c_void_p]
        L.ntsm_oracle_mc_destroy.argtypes = [C.c_void_p]
        L.ntsm_oracle_mc_insert.argtypes = [C.c_void_p, C.c_uint, C.c_uint64, C.c_uint, C.c_void_p]
        L.ntsm_oracle_mc_matrix.restype = C.POINTER(C.c_uint8)
        L.ntsm_oracle_mc_matrix.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]

    # -- primitives ---------------------------------------------------------------
    def nt4(self, b):
        return self.L.ntsm_oracle_nt4(b)

    def hash64(self, x, k=19):
        return self.L.ntsm_oracle_hash64(x, k)

    def iter(self, seq: bytes, k=19):
        """[(pos, hash, fw, rv), ...] as the reference iterator yields them."""
        cap = max(1, len(seq))
        h = np.zeros(cap, np.uint64); p = np.zeros(cap, np.uint64)
        fw = np.zeros(cap, np.uint64); rv = np.zeros(cap, np.uint64)
        n = self.L.ntsm_oracle_iter(seq, len(seq), k, h.ctypes.data_as(u64p), p.ctypes.data_as(u64p),
                                    fw.ctypes.data_as(u64p), rv.ctypes.data_as(u64p), cap)
        return [(int(p[i]), int(h[i]), int(fw[i]), int(rv[i])) for i in range(n)]

    def read_records(self, path):
        """All (name, seq) records + the terminating negative code, via the kseq restatement."""
        r = self.L.ntsm_oracle_reader_open(os.fsencode(path))
        assert r, path
        out = []
        s = C.c_char_p(); nm = C.c_char_p()
        while True:
            l = self.L.ntsm_oracle_reader_next(r, C.byref(s), C.byref(nm))
            if l < 0:
                break
            out.append((nm.value, C.string_at(s, l)))
        self.L.ntsm_oracle_reader_close(r)
        return out, l

    def vcf_run(self, sites, ref, vcf, prefix, k=19, dupes=0, multi=20, window=31):
        """VCFConvert + MultiCount restated, one thread; writes <prefix>_matrix.tsv, _center.txt, _counts_<j>.txt,
        _mat.bin, _stderr.txt (the file layout of tools/ref_vcf_harness.cpp).  Returns 0, -134 (the reference
        dies), -2 (undefined upstream), -1 (I/O)."""
        return self.L.ntsm_oracle_vcf_run(os.fsencode(sites), os.fsencode(ref), os.fsencode(vcf), k, int(dupes), multi, window,
                                          os.fsencode(prefix))

    def multicount(self, sites_path, n_samples, k=19, dupes=False):
        return OracleMC(self, sites_path, n_samples, k, dupes)

    def fingerprint(self, sites_path, k=19, dupes=False, cov=0.0):
        return OracleFP(self, sites_path, k, dupes, cov)


class OracleFP:
    """FingerPrint restated (src/FingerPrint.hpp); see oracle/ntsm_oracle.h."""

    def __init__(self, o, sites_path, k, dupes, cov):
        self.o, self.L, self.k = o, o.L, k
        self.h = self.L.ntsm_oracle_fp_create(os.fsencode(sites_path), k, int(dupes), float(cov), None)
        if not self.h:
            raise FileNotFoundError(sites_path)

    def close(self):
        if self.h:
            self.L.ntsm_oracle_fp_destroy(self.h)
            self.h = None

    __del__ = close

    def insert(self, seq: bytes):
        self.L.ntsm_oracle_fp_insert(self.h, seq, len(seq))

    def insert_many(self, buf: np.ndarray, off: np.ndarray, threads=8):
        buf = np.ascontiguousarray(buf, np.uint8); off = np.ascontiguousarray(off, np.uint64)
        self.L.ntsm_oracle_fp_insert_many(self.h, buf.ctypes.data, off.ctypes.data, len(off) - 1, threads)

    def count_file(self, path):
        rc = self.L.ntsm_oracle_fp_count_file(self.h, os.fsencode(path))
        if rc:
            raise FileNotFoundError(path)

    @property
    def total_kmers(self): return self.L.ntsm_oracle_fp_total_kmers(self.h)
    @property
    def total_counts(self): return self.L.ntsm_oracle_fp_total_counts(self.h)
    @property
    def total_bases(self): return self.L.ntsm_oracle_fp_total_bases(self.h)
    @property
    def max_counts(self): return self.L.ntsm_oracle_fp_max_counts(self.h)
    @property
    def early_term(self): return bool(self.L.ntsm_oracle_fp_early_term(self.h))
    @property
    def table_size(self): return self.L.ntsm_oracle_fp_table_size(self.h)
    @property
    def n_sites(self): return self.L.ntsm_oracle_fp_n_sites(self.h)

    def rows(self):
        n = self.n_sites
        a = [np.zeros(n, np.uint32) for _ in range(6)]
        rc = self.L.ntsm_oracle_fp_rows(self.h, *[x.ctypes.data for x in a])
        if rc:
            raise KeyError("Couldn't find key.")
        return a

    def lists(self):
        n = self.L.ntsm_oracle_fp_n_listed(self.h)
        hs = np.zeros(n, np.uint64); off = np.zeros(2 * self.n_sites + 1, np.uint32); cn = np.zeros(n, np.uint32)
        self.L.ntsm_oracle_fp_lists(self.h, hs.ctypes.data, off.ctypes.data, cn.ctypes.data)
        return hs, off, cn

    def names(self):
        return [self.L.ntsm_oracle_fp_site_name(self.h, i).decode() for i in range(self.n_sites)]

    def counts_text(self):
        """The counts file exactly as printOptionalHeader + printCountsMax write it."""
        mr, mv, sr, sv, nr, nv = self.rows()
        out = io.StringIO()
        out.write("#@TK\t%d\n#@KS\t%d" % (self.total_kmers, self.k))
        out.write("\n#locusID\tcountAT\tcountCG\tsumAT\tsumCG\tdistinctAT\tdistinctCG\n")
        for i, nm in enumerate(self.names()):
            out.write("%s\t%d\t%d\t%d\t%d\t%d\t%d\n" % (nm, mr[i], mv[i], sr[i], sv[i], nr[i], nv[i]))
        return out.getvalue()

    def summary(self):
        b = C.create_string_buffer(2048)
        self.L.ntsm_oracle_fp_summary(self.h, b, 2048)
        return b.value.decode()


class OracleMC:
    """MultiCount restated (src/MultiCount.hpp:43-70); see oracle/ntsm_oracle.h."""

    def __init__(self, o, sites_path, n_samples, k, dupes):
        self.L, self.n_samples = o.L, n_samples
        self.h = self.L.ntsm_oracle_mc_create(os.fsencode(sites_path), k, int(dupes), n_samples, None)
        if not self.h:
            raise FileNotFoundError(sites_path)

    def close(self):
        if self.h:
            self.L.ntsm_oracle_mc_destroy(self.h)
            self.h = None

    __del__ = close

    def insert(self, sample, hash_value, multi):
        self.L.ntsm_oracle_mc_insert(self.h, sample, hash_value, multi, None)

    def matrix(self):
        n = C.c_uint64()
        p = self.L.ntsm_oracle_mc_matrix(self.h, C.byref(n))
        a = np.ctypeslib.as_array(p, (n.value,)).copy() if n.value else np.zeros(0, np.uint8)
        return a.reshape(self.n_samples, -1) if self.n_samples else a.reshape(0, 0)


def load(path):
    return Oracle(path)

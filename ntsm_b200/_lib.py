"""ctypes binding of include/ntsm_b200.h.  Fails loudly if the library has not been built."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libntsm_b200.so")


class NtsmError(RuntimeError):
    def __init__(self, code, text):
        super().__init__("ntsm_b200 error %d: %s" % (code, text))
        self.code = code


class Cfg(C.Structure):
    _fields_ = [("k", C.c_uint32), ("device", C.c_int32), ("n_buffers", C.c_uint32), ("reserved", C.c_uint32),
                ("batch_bases", C.c_uint64), ("max_counts", C.c_uint64)]


_P = C.c_void_p
_SIGS = {
    # name: (restype, argtypes)
    "ntsm_version": (C.c_char_p, []),
    "ntsm_device_count": (C.c_int, []),
    "ntsm_device_warmup": (C.c_int, [C.c_int]),
    "ntsm_last_error": (C.c_char_p, [_P]),
    "ntsm_nt4": (C.c_uint32, [C.c_uint8]),
    "ntsm_hash64": (C.c_uint64, [C.c_uint64, C.c_uint32]),
    "ntsm_hash64_inv": (C.c_uint64, [C.c_uint64, C.c_uint32]),
    "ntsm_sites_load": (C.c_int, [C.POINTER(_P), C.c_char_p, C.c_uint32, C.c_int]),
    "ntsm_sites_free": (None, [_P]),
    "ntsm_sites_k": (C.c_uint32, [_P]),
    "ntsm_sites_n_sites": (C.c_uint32, [_P]),
    "ntsm_sites_n_kmers": (C.c_uint32, [_P]),
    "ntsm_sites_table_size": (C.c_uint64, [_P]),
    "ntsm_sites_hashes": (C.POINTER(C.c_uint64), [_P]),
    "ntsm_sites_allele_off": (C.POINTER(C.c_uint32), [_P]),
    "ntsm_sites_erased": (C.POINTER(C.c_uint8), [_P]),
    "ntsm_sites_name": (C.c_char_p, [_P, C.c_uint32]),
    "ntsm_sites_n_warnings": (C.c_uint32, [_P]),
    "ntsm_sites_warning": (C.c_char_p, [_P, C.c_uint32]),
    "ntsm_sites_printable": (C.c_int, [_P]),
    "ntsm_sites_max_counts": (C.c_uint64, [_P, C.c_double]),
    "ntsm_ctx_create": (C.c_int, [C.POINTER(_P), C.POINTER(Cfg)]),
    "ntsm_ctx_destroy": (None, [_P]),
    "ntsm_load_sites": (C.c_int, [_P, _P, _P, C.c_uint32, _P, C.c_uint32]),
    "ntsm_load_siteset": (C.c_int, [_P, _P]),
    "ntsm_ctx_set_option": (C.c_int, [_P, C.c_char_p, C.c_int]),
    "ntsm_host_register": (C.c_int, [_P, C.c_uint64]),
    "ntsm_host_unregister": (C.c_int, [_P]),
    "ntsm_group_finalize": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P, _P]),
    "ntsm_ctx_pcie_bytes": (None, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "ntsm_ctx_l2_window": (C.c_int, [_P]),
    "ntsm_ctx_probe_bytes": (C.c_uint64, [_P]),
    "ntsm_padded_positions": (C.c_uint64, [C.c_uint64]),
    "ntsm_acquire_batch": (C.c_int, [_P, C.POINTER(_P)]),
    "ntsm_batch_append": (C.c_int, [_P, C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]),
    "ntsm_batch_positions": (C.c_uint64, [_P]),
    "ntsm_batch_bases": (C.c_uint64, [_P]),
    "ntsm_batch_reads": (C.c_uint64, [_P]),
    "ntsm_submit_batch": (C.c_int, [_P, _P]),
    "ntsm_release_batch": (C.c_int, [_P, _P]),
    "ntsm_pack_isa": (C.c_char_p, [C.c_char_p]),
    "ntsm_pack_reads": (C.c_uint64, [_P, _P, C.c_uint64, _P, _P, _P]),
    "ntsm_pack_reads2": (C.c_uint64, [_P, _P, C.c_uint64, _P, _P, _P, C.c_int]),
    "ntsm_count_packed_device": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64, _P]),
    "ntsm_count_packed_host": (C.c_int, [_P, _P, _P, C.c_uint64, C.c_uint64]),
    "ntsm_reset_counts_async": (C.c_int, [_P]),
    "ntsm_set_stream": (C.c_int, [_P, _P]),
    "ntsm_reduce_async": (C.c_int, [_P]),
    "ntsm_insert_count": (C.c_int, [_P, C.c_char_p, C.c_uint64]),
    "ntsm_flush": (C.c_int, [_P]),
    "ntsm_insert_reads": (C.c_int, [_P, C.c_uint32, _P, _P, C.c_uint64, C.c_uint32]),
    "ntsm_insert_reads_fixed": (C.c_int, [_P, C.c_uint32, _P, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32]),
    "ntsm_poll_totals": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_int)]),
    "ntsm_sync": (C.c_int, [_P]),
    "ntsm_reset_counts": (C.c_int, [_P]),
    "ntsm_nccl_unique_id": (C.c_int, [_P]),
    "ntsm_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "ntsm_allreduce": (C.c_int, [_P]),
    "ntsm_finalize": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "ntsm_get_counts": (C.c_int, [_P, _P]),
    "ntsm_get_totals": (C.c_int, [_P, _P]),
    "ntsm_add_counts": (C.c_int, [_P, _P, _P]),
    "ntsm_counts_save": (C.c_int, [_P, _P, C.c_char_p]),
    "ntsm_counts_load_add": (C.c_int, [_P, _P, C.c_char_p]),
    "ntsm_sites_covered": (C.c_uint32, [_P, _P, C.c_uint32]),
    "ntsm_format_counts": (C.c_int64, [_P, _P, _P, _P, _P, C.c_uint64, _P, C.c_size_t]),
    "ntsm_format_summary": (C.c_int64, [_P, _P, C.c_uint32, _P, C.c_size_t]),
    "ntsm_ctx_launches": (C.c_uint64, [_P]),
    "ntsm_ctx_kernel_name": (C.c_char_p, [_P]),
    "ntsm_ctx_filter_bits": (C.c_uint32, [_P]),
    "ntsm_ctx_table_capacity": (C.c_uint32, [_P]),
    "ntsm_reader_open": (C.c_int, [C.POINTER(_P), C.c_char_p]),
    "ntsm_reader_next": (C.c_int64, [_P, C.POINTER(C.c_char_p)]),
    "ntsm_reader_name": (C.c_char_p, [_P]),
    "ntsm_reader_close": (None, [_P]),
    "ntsm_reader_open2": (C.c_int, [C.POINTER(_P), C.c_char_p, C.c_int]),
    "ntsm_scan_isa": (C.c_char_p, [C.c_char_p]),
    "ntsm_gz_open": (C.c_int, [C.POINTER(_P), C.c_char_p, C.c_int]),
    "ntsm_gz_read": (C.c_int, [_P, C.c_void_p, C.c_uint]),
    "ntsm_gz_mode": (C.c_char_p, [_P]),
    "ntsm_gz_fell_back": (C.c_int, [_P]),
    "ntsm_gz_parallel_chunks": (C.c_uint64, [_P]),
    "ntsm_gz_close": (None, [_P]),
    "ntsm_crc32": (C.c_uint32, [C.c_uint32, C.c_void_p, C.c_uint64]),
    "ntsm_count_files": (C.c_int, [_P, C.c_uint32, _P, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_int)]),
    "ntsm_main": (C.c_int, [C.c_int, _P]),
    # multi-sample matrix path (MultiCount / VCFConvert)
    "ntsm_multi_create": (C.c_int, [C.POINTER(_P), _P, C.c_uint32]),
    "ntsm_multi_destroy": (None, [_P]),
    "ntsm_multi_n_samples": (C.c_uint32, [_P]),
    "ntsm_multi_insert_count": (C.c_int, [_P, C.c_uint32, C.c_uint64, C.c_uint32]),
    "ntsm_multi_insert_windows": (C.c_int, [_P, _P, C.c_uint32, _P, _P, C.c_uint32, C.c_uint32]),
    "ntsm_multi_n_warnings": (C.c_uint64, [_P]),
    "ntsm_multi_warnings_text": (C.c_int64, [_P, _P, C.c_size_t]),
    "ntsm_multi_get_matrix": (C.c_int, [_P, _P]),
    "ntsm_multi_kernel_ms": (None, [_P, _P, C.POINTER(C.c_uint64)]),
    "ntsm_multi_counts_max": (C.c_int, [_P, C.c_uint32, _P, _P, _P, _P]),
    "ntsm_multi_format_counts": (C.c_int64, [_P, _P, C.c_uint32, _P, C.c_size_t]),
    "ntsm_multi_norm_matrix": (C.c_int, [_P, _P, _P]),
    "ntsm_multi_write_norm_matrix": (C.c_int, [_P, _P, _P, C.c_char_p, C.c_char_p, C.c_uint32]),
    "ntsm_vcf_convert": (C.c_int, [C.POINTER(_P), _P, _P, C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]),
    "ntsm_vcf_destroy": (None, [_P]),
    "ntsm_vcf_parse": (C.c_int, [C.POINTER(_P), C.c_char_p, C.c_char_p, C.c_uint32, C.c_uint32, C.c_int]),
    "ntsm_vcf_lines_free": (None, [_P]),
    "ntsm_vcf_lines_n_samples": (C.c_uint32, [_P]),
    "ntsm_vcf_lines_sample_id": (C.c_char_p, [_P, C.c_uint32]),
    "ntsm_vcf_lines_count": (C.c_uint64, [_P]),
    "ntsm_vcf_lines_wstride": (C.c_uint32, [_P]),
    "ntsm_vcf_lines_windows": (_P, [_P]),
    "ntsm_vcf_lines_lens": (_P, [_P]),
    "ntsm_vcf_lines_genotypes": (_P, [_P]),
    "ntsm_vcf_stream_chunk": (C.c_uint64, [C.c_uint64]),
    "ntsm_vcf_genotype_isa": (C.c_int, [C.c_int]),
    "ntsm_vcf_multi": (_P, [_P]),
    "ntsm_vcf_n_samples": (C.c_uint32, [_P]),
    "ntsm_vcf_sample_id": (C.c_char_p, [_P, C.c_uint32]),
    "ntsm_vcf_lines_counted": (C.c_uint64, [_P]),
    "ntsm_vcf_output_matrix": (C.c_int, [_P, C.c_char_p]),
    "ntsm_vcf_output_counts": (C.c_int, [_P, C.c_char_p]),
    "ntsm_vcf_main": (C.c_int, [C.c_int, _P]),
}

_lib = None


def lib_path():
    return _LIB_PATH


def _preload_bundled_nccl():
    """libntsm_b200.so needs libnccl.so.2.  torch ships a newer one than the system's and refuses to
    start on the older: whichever of the two is mapped first serves both, so map torch's first (it
    is what the library runs on whenever torch was imported before it, i.e. in bench.py and the
    multi-rank tests)."""
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for d in (spec.submodule_search_locations if spec else []):
            so = os.path.join(d, "lib", "libnccl.so.2")
            if os.path.exists(so):
                C.CDLL(so, mode=C.RTLD_GLOBAL)
                return so
    except Exception:
        pass
    return None


def lib():
    """The loaded library.  No fallback: a missing build is an error, not a slow path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C ntsm_b200/csrc` (there is no CPU fallback)" % _LIB_PATH)
        _preload_bundled_nccl()
        L = C.CDLL(_LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in _SIGS.items():
            f = getattr(L, name)          # AttributeError here = header/library mismatch
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def check(rc, ctx=None):
    if rc < 0:
        raise NtsmError(rc, (lib().ntsm_last_error(ctx) or b"").decode(errors="replace"))
    return rc

"""ntsm_b200 -- B200-native counting hot path of ntsmCount (JustinChu/ntsm).

The product is the C-ABI library `ntsm_b200/lib/libntsm_b200.so` (CUDA sm_100a kernels + host
ingest) and the `ntsm_b200/bin/ntsmCount` command line.  This package is only the Python view of
that ABI: a ctypes binding (`_lib`) and a mirror of the reference's FingerPrint object
(`FingerPrint`, src/FingerPrint.hpp; `MultiCount` / `VCFConvert` for the multi-sample matrix path) so tests and the benchmark read like the reference's own
call sequence.  There is no CPU fallback: importing works anywhere, counting needs a GPU.
"""
from ._lib import NtsmError, lib, lib_path  # noqa: F401
from .fingerprint import FingerPrint, SiteSet, pack_reads  # noqa: F401
from .multicount import MultiCount, VCFConvert  # noqa: F401

__all__ = ["FingerPrint", "SiteSet", "pack_reads", "MultiCount", "VCFConvert", "lib", "lib_path", "NtsmError"]

"""popbam_b200 -- B200-native implementation of POPBAM's per-window statistics path.

The product is the C-ABI shared library built from popbam_b200/csrc (include/popbam_b200.h) and the
`popbam` command line on top of it.  This Python package is a thin ctypes binding used by the tests and
bench.py; it contains no compute and fails loudly when the CUDA library has not been built.
"""
from .capi import (AN, FLAG, Batch, Context, Params, PrintOpts, Result, build, lib, lib_path)  # noqa: F401

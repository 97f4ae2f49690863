"""b200-radiosity: B200-native progressive-refinement shooting loop (hw-radiosity's hot path).

The product is the C-ABI library `libradk.so` (include/radk.h): hand-written sm_100a CUDA kernels
plus the C++ host front end, and the `rad` command-line driver.  This Python package is only
plumbing for tests and benchmarks: a ctypes binding (`api`) and the multi-GPU launcher glue.
"""
from .api import (Context, Model, RadkError, build_library, default_opts, library_path, load_library,  # noqa: F401
                  STATS_FIELDS, TRI_DTYPE, NODE_DTYPE)

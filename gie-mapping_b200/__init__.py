"""gie-mapping_b200 — B200-native OGM + incremental EDT engine behind the GIE-mapping operator surface.

The product is the CUDA library `libgie_b200.so` (csrc/, C ABI in include/gie_b200.h).  This package is the thin
Python host mirror used by tests and bench.py; it binds the C ABI with ctypes and FAILS LOUDLY when the library is
missing — there is no CPU fallback.
"""
from .engine import (  # noqa: F401
    GieError,
    LocMap,
    GlbHashMap,
    Mapper,
    load_library,
    library_path,
    STAGE_NAMES,
    ARR_RAY_COUNT, ARR_INST_TYPE, ARR_GLB_TYPE, ARR_EDT, ARR_AUX, ARR_COC_AUX, ARR_PAIR,
)
from . import scenes  # noqa: F401
from . import replay_io  # noqa: F401

__all__ = ["GieError", "LocMap", "GlbHashMap", "Mapper", "load_library", "library_path", "scenes", "STAGE_NAMES"]

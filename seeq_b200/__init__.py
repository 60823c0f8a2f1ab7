"""seeq_b200 -- B200-native drop-in for seeq's per-line Levenshtein matching path.

The product is ``libseeq_b200.so`` (hand-written sm_100a kernels behind the
unchanged libseeq C API, see include/*.h and DESIGN.md).  This Python package
only holds the build recipe, a ctypes binding used by the tests and bench.py,
and the multi-GPU sharding driver.
"""
from . import build  # noqa: F401

__all__ = ["build", "binding", "shard"]

"""Import shim for the *live* reference (test harness only, never product code): thin alias of oracle/reference.py, which holds
the stubs for pygame / gymnasium / rvo2 / socialforce / matplotlib and finds the reference at /root/reference (build container)
or under oracle/_ref (staged copy on the GPU box).  Only tests/golden/make_golden.py and the live cross-check tests use it."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import reference as _reference  # noqa: E402

REFERENCE_ROOT = _reference.root()
reference_available = _reference.available
install = _reference.install

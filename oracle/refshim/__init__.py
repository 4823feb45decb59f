"""Compat loader that imports the UNMODIFIED reference from /root/reference.

Test infrastructure only (see oracle/README.md).  Used in the build container
to pin the oracle and to generate tests/golden/* fixtures.  Never imported by
the product package, bench.py's GPU arm, or anything that runs on the GPU box
(/root/reference does not exist there).
"""
from .loader import load_reference, reference_available  # noqa: F401

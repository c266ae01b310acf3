"""Import the unmodified reference for the tests.  TEST HELPER ONLY.

Thin wrapper over ``oracle/ref_env.py``: the reference is found at ``$SENAS_REF``, ``/root/reference`` (build
container) or ``oracle/_ref`` (staged byte-for-byte by ``oracle/make_ref.py``; what the GPU box has), with the stub
set of SURVEY.md section 8c for the third-party modules that are not installed.  Callers are skipped when
``available()`` is False.
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import ref_env  # noqa: E402

REF = ref_env.root()


def available():
    return ref_env.available()


def load():
    """Returns the reference's (search.cell, search.senas_search, utils.operations) modules."""
    return ref_env.load()

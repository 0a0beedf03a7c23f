"""Error conventions of the reference kept at the Python layer (error.py:12-28 there): KLTError reports and terminates
the process with exit status 1, KLTWarning reports and carries on.  Nothing of this crosses the C ABI, which returns
status codes instead."""
from __future__ import print_function
import sys


def _report(message):
    print(message)


def KLTError(err):
    """Fatal: print the message and leave the interpreter with status 1 (SystemExit, like the reference's exit(1))."""
    _report(err)
    sys.exit(1)


def KLTWarning(err):
    """Non-fatal: print the message and continue."""
    _report(err)

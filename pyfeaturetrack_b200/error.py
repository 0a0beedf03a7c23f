"""error.py of the reference (error.py:12-28): KLTError prints and exits, KLTWarning prints."""
from __future__ import print_function


def KLTError(err):
    print(err)
    exit(1)


def KLTWarning(err):
    print(err)

"""`natsort.natsorted` / `realsorted` / `humansorted`: natural ordering of strings with embedded
numbers (reference use: cluster_helpers.py:8 import; pixel_cluster_utils.py:13)."""
import re

_NUM = re.compile(r"(\d+)")


def _key(s):
    parts = _NUM.split(str(s))
    return [(0, int(p)) if i % 2 else (1, p.lower(), p) for i, p in enumerate(parts)]


def natsort_keygen(key=None, alg=0):
    return (lambda v: _key(key(v))) if key else _key


def natsorted(seq, key=None, reverse=False, alg=0):
    return sorted(seq, key=natsort_keygen(key), reverse=reverse)


realsorted = humansorted = os_sorted = natsorted


class ns:  # algorithm flags accepted and ignored
    DEFAULT = INT = REAL = IGNORECASE = PATH = 0

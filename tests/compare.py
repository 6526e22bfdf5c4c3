"""Array-by-array comparison of two MCED dumps (tests/mceio.py).  Test infrastructure."""
import re
import sys

import numpy as np

from mceio import read_dump

_SKIP = re.compile(r"/ms$|^header$")


def _sorted_tables(d, pre):
    """Per-term (keys, G) with keys sorted -- already sorted in every producer -- and encB as sorted sets."""
    cells = d[pre + "/cells"]
    off = np.concatenate([[0], np.cumsum(cells)])
    return cells, off


def compare_dumps(a, b, float_rtol=0.0, float_names_rtol=None, verbose=False, skip=None, encB_as_set=True):
    """Returns a list of mismatch strings (empty == equal). Integer arrays must match exactly; float arrays
    within float_rtol (relative to the max |value| of the array; 0.0 = bit-exact)."""
    problems = []
    float_names_rtol = float_names_rtol or {}
    names = [n for n in a.keys() if not _SKIP.search(n) and not (skip and skip(n))]
    for n in names:
        if n not in b:
            problems.append("missing in B: " + n)
            continue
        x, y = a[n], b[n]
        if x.shape != y.shape:
            problems.append("%s: shape %s vs %s" % (n, x.shape, y.shape))
            continue
        if n.endswith("/encB") and encB_as_set:
            pre = n[: -len("/encB")]
            cells, off = _sorted_tables(a, pre)
            bad = 0
            for i in range(len(cells)):
                if not np.array_equal(np.sort(x[off[i] : off[i + 1]]), np.sort(y[off[i] : off[i + 1]])):
                    bad += 1
            if bad:
                problems.append("%s: %d terms differ as key sets" % (n, bad))
            continue
        if x.dtype.kind in "iub":
            if not np.array_equal(x, y):
                idx = np.argwhere(x != y)
                problems.append("%s: %d/%d integer entries differ, first at %s: %s vs %s"
                                % (n, len(idx), x.size, idx[0], x[tuple(idx[0])], y[tuple(idx[0])]))
            continue
        rtol = float_rtol
        for pat, r in float_names_rtol.items():
            if re.search(pat, n):
                rtol = r
        if rtol == 0.0:
            same = (x == y) | (np.isnan(x) & np.isnan(y)) if x.dtype.kind == "f" else (x == y) | (np.isnan(x.real) & np.isnan(y.real))
            if not same.all():
                idx = np.argwhere(~same)
                scale = np.max(np.abs(x)) if x.size else 0.0
                err = np.max(np.abs(x - y)[~same])
                problems.append("%s: %d/%d float entries not bit-equal (max abs diff %.3e, scale %.3e), first at %s"
                                % (n, len(idx), x.size, err, scale, idx[0]))
        else:
            scale = max(np.max(np.abs(x)) if x.size else 0.0, 1e-300)
            err = np.max(np.abs(x - y)) if x.size else 0.0
            if not (err <= rtol * scale):
                problems.append("%s: max abs diff %.3e > %.1e * %.3e" % (n, err, rtol, scale))
    for n in b.keys():
        if n not in a and not _SKIP.search(n) and not (skip and skip(n)):
            problems.append("missing in A: " + n)
    if verbose:
        print("compared %d arrays, %d problems" % (len(names), len(problems)))
    return problems


if __name__ == "__main__":
    A = read_dump(sys.argv[1])
    B = read_dump(sys.argv[2])
    rtol = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
    probs = compare_dumps(A, B, float_rtol=rtol, verbose=True)
    for p in probs[:60]:
        print("  ", p)
    sys.exit(1 if probs else 0)

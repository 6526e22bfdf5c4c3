"""Readers/writers for the scenario (MCES) and dump (MCED) files defined in oracle/mce_io.h.

Test infrastructure only: scenarios are open-loop recordings of the arguments passed to
CauchyEstimator::step() (reference include/cauchy_estimator.hpp:1211); dumps are flat lists of
named arrays written by oracle/ref_run.cpp, oracle/mce_oracle_run.c and tests/helpers for the GPU path.
"""
import struct
from dataclasses import dataclass, field

import numpy as np

MCES_MAGIC = 0x5345434D
MCED_MAGIC = 0x4445434D
SHIFT_NONE, SHIFT_OWN_MEAN, SHIFT_EXPLICIT = 0, 1, 2

_DT = {0: np.float64, 1: np.int32, 2: np.uint32, 3: np.uint8, 4: np.int8, 5: np.complex128}


@dataclass
class StepRecord:
    msmt: float
    gamma: float
    Phi: np.ndarray
    Gamma: np.ndarray
    beta: np.ndarray
    H: np.ndarray
    B: np.ndarray = None
    u: np.ndarray = None
    shift_kind: int = SHIFT_NONE
    delta: np.ndarray = None


@dataclass
class Scenario:
    d: int
    cmcc: int
    pncc: int
    p: int
    steps: int
    tr_order: list
    root_point: np.ndarray
    b_pert: np.ndarray
    A0: np.ndarray
    p0: np.ndarray
    b0: np.ndarray
    rec: list = field(default_factory=list)

    @property
    def max_shape(self):
        # reference include/cauchy_estimator.hpp:97
        return (self.steps - 1) * self.pncc + self.d if self.d > 1 else self.d + self.pncc


def write_scenario(path, s: Scenario):
    d = s.d
    with open(path, "wb") as f:
        f.write(struct.pack("<I", MCES_MAGIC))
        f.write(struct.pack("<7i", 1, s.d, s.cmcc, s.pncc, s.p, s.steps, len(s.rec)))
        order = list(s.tr_order) + list(range(len(s.tr_order), 12))
        f.write(struct.pack("<12i", *order[:12]))
        f.write(np.asarray(s.root_point, np.float64)[:d].tobytes())
        ms = s.max_shape
        f.write(struct.pack("<i", ms))
        bp = np.zeros(ms)
        bp[: min(ms, len(s.b_pert))] = np.asarray(s.b_pert, np.float64)[:ms]
        f.write(bp.tobytes())
        f.write(np.asarray(s.A0, np.float64).reshape(d * d).tobytes())
        f.write(np.asarray(s.p0, np.float64).reshape(d).tobytes())
        f.write(np.asarray(s.b0, np.float64).reshape(d).tobytes())
        for r in s.rec:
            f.write(struct.pack("<2d", r.msmt, r.gamma))
            f.write(np.asarray(r.Phi, np.float64).reshape(d * d).tobytes())
            f.write(np.asarray(r.Gamma, np.float64).reshape(d * s.pncc).tobytes())
            f.write(np.asarray(r.beta, np.float64).reshape(s.pncc).tobytes())
            f.write(np.asarray(r.H, np.float64).reshape(d).tobytes())
            has_bu = r.B is not None and r.u is not None and s.cmcc > 0
            f.write(struct.pack("<i", int(has_bu)))
            if has_bu:
                f.write(np.asarray(r.B, np.float64).reshape(d * s.cmcc).tobytes())
                f.write(np.asarray(r.u, np.float64).reshape(s.cmcc).tobytes())
            f.write(struct.pack("<i", r.shift_kind))
            delta = np.zeros(d) if r.delta is None else np.asarray(r.delta, np.float64).reshape(d)
            f.write(delta.tobytes())


def read_scenario(path) -> Scenario:
    with open(path, "rb") as f:
        buf = f.read()
    off = 0

    def ints(n):
        nonlocal off
        v = struct.unpack_from("<%di" % n, buf, off)
        off += 4 * n
        return list(v)

    def dbl(n):
        nonlocal off
        v = np.frombuffer(buf, np.float64, n, off).copy()
        off += 8 * n
        return v

    (magic,) = struct.unpack_from("<I", buf, off)
    off += 4
    assert magic == MCES_MAGIC, "not a MCES file"
    _, d, cmcc, pncc, p, steps, nrec = ints(7)
    order = ints(12)
    root_point = dbl(d)
    (ms,) = ints(1)
    b_pert = dbl(ms)
    A0 = dbl(d * d).reshape(d, d)
    p0 = dbl(d)
    b0 = dbl(d)
    s = Scenario(d, cmcc, pncc, p, steps, order, root_point, b_pert, A0, p0, b0)
    for _ in range(nrec):
        msmt, gamma = dbl(2)
        Phi = dbl(d * d).reshape(d, d)
        Gamma = dbl(d * pncc).reshape(d, pncc)
        beta = dbl(pncc)
        H = dbl(d)
        (has_bu,) = ints(1)
        B = u = None
        if has_bu:
            B = dbl(d * cmcc).reshape(d, cmcc)
            u = dbl(cmcc)
        (kind,) = ints(1)
        delta = dbl(d)
        s.rec.append(StepRecord(float(msmt), float(gamma), Phi, Gamma, beta, H, B, u, kind, delta))
    return s


def read_dump(path) -> dict:
    with open(path, "rb") as f:
        buf = f.read()
    (magic,) = struct.unpack_from("<I", buf, 0)
    assert magic == MCED_MAGIC, "not a MCED file"
    off = 4
    out = {}
    n = len(buf)
    while off < n:
        (nl,) = struct.unpack_from("<I", buf, off)
        off += 4
        name = buf[off : off + nl].decode()
        off += nl
        dt, nd = struct.unpack_from("<II", buf, off)
        off += 8
        dims = struct.unpack_from("<%dQ" % nd, buf, off)
        off += 8 * nd
        cnt = int(np.prod(dims)) if nd else 1
        dtype = np.dtype(_DT[dt])
        arr = np.frombuffer(buf, dtype, cnt, off).reshape(dims).copy() if cnt else np.zeros(dims, dtype)
        off += cnt * dtype.itemsize
        out[name] = arr
    return out


def write_dump(path, arrays: dict):
    inv = {np.dtype(v): k for k, v in _DT.items()}
    with open(path, "wb") as f:
        f.write(struct.pack("<I", MCED_MAGIC))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<II", inv[a.dtype], a.ndim))
            f.write(struct.pack("<%dQ" % a.ndim, *a.shape))
            f.write(a.tobytes())


def mix64(k):
    """Vectorised twin of mix64() in oracle/ref_run.cpp (key digest)."""
    k = np.asarray(k, np.uint64)
    with np.errstate(over="ignore"):
        x = (k + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(29)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(32)
    return x


def key_digest(cells, keys):
    """Returns the 8 x u32 digest ref_run writes for one shape: n, sum(cells), xor / rank-weighted sum of term hashes."""
    cells = np.asarray(cells, np.int64)
    n = len(cells)
    keys = np.asarray(keys, np.uint64)
    mh = mix64(keys)
    th = np.zeros(n, np.uint64)
    if len(keys):
        starts = np.concatenate([[0], np.cumsum(cells)[:-1]])
        nz = cells > 0
        with np.errstate(over="ignore"):
            red = np.add.reduceat(mh, starts[nz]) if nz.any() else np.zeros(0, np.uint64)
        th[nz] = red
    hx = np.bitwise_xor.reduce(th) if n else np.uint64(0)
    with np.errstate(over="ignore"):
        hs = np.sum(th * np.arange(1, n + 1, dtype=np.uint64), dtype=np.uint64) if n else np.uint64(0)
    sc = int(cells.sum())
    hx = int(hx)
    hs = int(hs)
    return np.array(
        [n, 0, sc & 0xFFFFFFFF, sc >> 32, hx & 0xFFFFFFFF, hx >> 32, hs & 0xFFFFFFFF, hs >> 32], np.uint32
    )

"""The moment kernel (software-pipelined producers + two accumulator lanes, KMomentsSerial) must equal the reference's dependent
addition chain bit for bit (cauchy_estimator.hpp:307-338 adds every term's contribution in order), whatever the data does: ties,
cancellation, binade walks, subnormals, infinities, tile boundaries.  CPU: the emulated kernel body; GPU: the product."""
import ctypes as ct

import numpy as np
import pytest

from harness import Session, load_emu, load_product
from mceio import read_scenario
from moment_cases import cases, serial_sum
import os
from harness import ROOT


def _check(lib):
    sc = read_scenario(os.path.join(ROOT, "tests", "golden", "lti3.mces"))
    s = Session(lib, sc)
    dp = ct.POINTER(ct.c_double)
    try:
        for name, a in cases().items():
            for comp in (0, 1):                      # the sequence as the real part, then as the imaginary part (other part: another sequence)
                other = a[::-1].copy()
                g = np.zeros((len(a), 2)); g[:, comp] = a; g[:, 1 - comp] = other
                g = np.ascontiguousarray(g)
                out = np.zeros(2)
                rc = lib.mce_debug_moment_sums(s.h, len(a), 0, g.ctypes.data_as(dp), None, out.ctypes.data_as(dp))
                assert rc == 0
                for cc, seq in ((comp, a), (1 - comp, other)):
                    want = serial_sum(seq)
                    assert out[cc].tobytes() == np.float64(want).tobytes(), "%s: kernel %r != serial chain %r" % (name, out[cc], want)
        # a full set of quantities: fz, mean and covariance sums of random slots against numpy's own sequential chain
        rng = np.random.default_rng(5)
        n, d = 3000, 3
        g = rng.standard_normal((n, 2)); y = rng.standard_normal((n, d, 2))
        out = np.zeros(2 * (1 + d + d * d))
        assert lib.mce_debug_moment_sums(s.h, n, d, np.ascontiguousarray(g).ctypes.data_as(dp), np.ascontiguousarray(y).ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0

        def cm(a, b):        # complex product as (ac - bd, ad + bc): the finite branch of __muldc3
            return np.stack([a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1], a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0]], -1)
        assert out[0].tobytes() == np.float64(serial_sum(g[:, 0])).tobytes() and out[1].tobytes() == np.float64(serial_sum(g[:, 1])).tobytes()
        for j in range(d):
            w = cm(g, y[:, j])
            for cc in range(2):
                assert out[2 * (1 + j) + cc].tobytes() == np.float64(serial_sum(w[:, cc])).tobytes()
            for k in range(d):
                w2 = -cm(w, y[:, k])
                for cc in range(2):
                    assert out[2 * (1 + d + j * d + k) + cc].tobytes() == np.float64(serial_sum(w2[:, cc])).tobytes()
    finally:
        s.close()


def _check_scan(lib):
    """KSumScan (Re fz of a partitioned estimator): the scan of rounding maps equals the dependent chain on every adversarial sequence."""
    sc = read_scenario(os.path.join(ROOT, "tests", "golden", "lti3.mces"))
    s = Session(lib, sc)
    dp = ct.POINTER(ct.c_double)
    restarts, fast = {}, {}
    try:
        for name, a in cases().items():
            g = np.zeros((len(a), 2)); g[:, 0] = a; g[:, 1] = 12345.0
            g = np.ascontiguousarray(g)
            out = np.zeros(3)
            assert lib.mce_debug_sum_scan(s.h, len(a), g.ctypes.data_as(dp), out.ctypes.data_as(dp)) == 0
            want = serial_sum(a)
            assert out[0].tobytes() == np.float64(want).tobytes(), "%s: scan %r != serial chain %r" % (name, out[0], want)
            restarts[name] = int(out[1]); fast[name] = int(out[2])
    finally:
        s.close()
    assert restarts["long_positive"] < 200 and restarts["all_zero"] == 0, restarts      # the scan path really carries the friendly chains
    assert 3 <= restarts["binade_walk"] < 40, restarts                                   # ... and the literal loop (in doubling runs) the hostile ones
    # the tiled path: 300 000 friendly addends are 37 tiles, nearly all applied from their summaries; hostile data falls back and is still exact
    assert fast["long_positive"] >= 30 and fast["tiles_ties"] >= 14, fast
    assert fast["tiles_big_offsets"] == 0 and fast["tiles_guess_off"] == 0 and fast["tiles_inf_inside"] < 10, fast      # wrong guesses cost time, never bits


def test_sum_scan_equals_the_serial_chain_emulated():
    _check_scan(load_emu())


@pytest.mark.gpu
def test_sum_scan_equals_the_serial_chain_gpu():
    _check_scan(load_product())


def test_moment_kernel_equals_the_serial_chain_emulated():
    _check(load_emu())


@pytest.mark.gpu
def test_moment_kernel_equals_the_serial_chain_gpu():
    _check(load_product())

"""Test harness: drives a C-ABI library (include/mce_b200.h) through a scenario and returns the same named
arrays that oracle/ref_run.cpp and oracle/mce_oracle_run.c dump, so results can be diffed with compare.py.

`load_product()` loads cauchyfriendly_b200/libmce_b200.so (CUDA, the product); `load_emu()` loads the
sequential emulation of the kernel bodies (tests/emu, CPU-only logic tests)."""
import ctypes as ct
import os
import subprocess

import numpy as np

from mceio import SHIFT_EXPLICIT, SHIFT_OWN_MEAN, key_digest, read_dump

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAXM = 32


import sys
sys.path.insert(0, ROOT)
from cauchyfriendly_b200._capi import MceMoments, MceOptions, MceStepStats, bind  # noqa: E402


def _dp(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_double)) if a is not None else None


def _emu_stale(so):
    """The emulation library is older than one of its sources (kernel headers, C ABI, the emulation backend)."""
    if not os.path.exists(so):
        return True
    t = os.path.getmtime(so)
    dirs = [os.path.join(ROOT, "cauchyfriendly_b200", "csrc"), os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "include")]
    return any(os.path.getmtime(os.path.join(d, f)) > t for d in dirs for f in os.listdir(d) if os.path.isfile(os.path.join(d, f)))


def load_emu(rebuild=False):
    """rebuild=True: make sure the library is up to date with its sources (a 20 s g++ run only when one of them changed)."""
    so = os.path.join(ROOT, "tests", "emu", "_build", "libmce_emu.so")
    if not os.path.exists(so) or (rebuild and _emu_stale(so)):
        subprocess.check_call([os.path.join(ROOT, "tests", "emu", "build.sh")])
    return bind(ct.CDLL(so))


def load_product():
    from cauchyfriendly_b200 import _capi
    return _capi.load()


class Session:
    def __init__(self, lib, sc, print_basic_info=False, device=-1, fast_moments=False, split=0, phase_timing=False, lean=False, early_scale=0):
        self.lib, self.sc = lib, sc
        o = MceOptions()
        lib.mce_default_options(ct.byref(o))
        for i in range(12):
            o.tr_search_order[i] = sc.tr_order[i]
        o.print_basic_info = int(print_basic_info)
        o.device = int(device)
        o.fast_moments = 1 if fast_moments else 0
        o.fast_moments_min_slots = int(fast_moments) if int(fast_moments) > 1 else 0      # fast_moments=N > 1: the option with its slot threshold set to N
        o.group_split_threshold = int(split)
        o.phase_timing = int(phase_timing)
        o.lean_group_kernel = int(lean)
        o.early_scale_min_slots = int(early_scale)
        self._keep = [np.ascontiguousarray(x, np.float64) for x in (sc.A0, sc.p0, sc.b0, sc.root_point, np.concatenate([sc.b_pert, np.zeros(MAXM)]))]
        self.h = lib.mce_create(sc.d, sc.cmcc, sc.pncc, sc.p, sc.steps, *[_dp(x) for x in self._keep], ct.byref(o))
        if not self.h:
            raise RuntimeError(lib.mce_last_error().decode())
        self.shape_range = lib.mce_shape_range(self.h)

    def close(self):
        if self.h:
            self.lib.mce_destroy(self.h)
            self.h = None

    def step(self, r):
        Phi = np.ascontiguousarray(r.Phi, np.float64)
        Gam = np.ascontiguousarray(r.Gamma, np.float64)
        beta = np.ascontiguousarray(r.beta, np.float64)
        H = np.ascontiguousarray(r.H, np.float64)
        B = np.ascontiguousarray(r.B, np.float64) if r.B is not None else None
        u = np.ascontiguousarray(r.u, np.float64) if r.u is not None else None
        rc = self.lib.mce_step(self.h, r.msmt, _dp(Phi), _dp(Gam), _dp(beta), _dp(H), r.gamma, _dp(B), _dp(u))
        if rc < 0:
            raise RuntimeError("mce_step failed (%d): %s" % (rc, self.lib.mce_last_error().decode()))
        return rc

    def moments(self):
        m = MceMoments()
        self.lib.mce_get_moments(self.h, ct.byref(m))
        return m

    def stats(self):
        s = MceStepStats()
        self.lib.mce_get_step_stats(self.h, ct.byref(s))
        return s

    def counts(self, after_muc):
        c = (ct.c_int * self.shape_range)()
        self.lib.mce_get_terms_per_shape(self.h, c, int(after_muc))
        return np.array(c[:], np.int32)

    def shift_b(self, delta, sign=-1.0):
        dl = np.ascontiguousarray(delta, np.float64)
        self.lib.mce_shift_b(self.h, _dp(dl), sign)

    def export_shape(self, m):
        d = self.sc.d
        n, tot = ct.c_int(0), ct.c_longlong(0)
        self.lib.mce_export_shape(self.h, m, ct.byref(n), ct.byref(tot), None, None, None, None, None, None)
        n, tot = n.value, tot.value
        A, p, b = np.zeros((n, m * d)), np.zeros((n, m)), np.zeros((n, d))
        cells, keys, G = np.zeros(n, np.int32), np.zeros(tot, np.uint32), np.zeros(tot, np.complex128)
        if n:
            n2, t2 = ct.c_int(0), ct.c_longlong(0)
            self.lib.mce_export_shape(self.h, m, ct.byref(n2), ct.byref(t2), _dp(A), _dp(p), _dp(b), cells.ctypes.data_as(ct.POINTER(ct.c_int)),
                                      keys.ctypes.data_as(ct.POINTER(ct.c_uint32)), G.view(np.float64).ctypes.data_as(ct.POINTER(ct.c_double)))
        return dict(A=A, p=p, b=b, cells=cells, keys=keys, G=G)

    def debug_muc_shape(self, m):
        d = self.sc.d
        n = ct.c_int(0)
        self.lib.mce_debug_muc_shape(self.h, m, ct.byref(n), None, None, None, None, None, None, None, None, None)
        n = n.value
        if n == 0:
            return None
        A, p, q, b, cd = np.zeros((n, m * d)), np.zeros((n, m)), np.zeros((n, m)), np.zeros((n, d)), np.zeros((n, 2))
        meta, cmap, F = np.zeros((n, 8), np.int32), np.zeros((n, MAXM), np.uint8), np.zeros(n, np.int32)
        csmap = np.zeros((n, MAXM), np.int8)
        self.lib.mce_debug_muc_shape(self.h, m, ct.byref(n := ct.c_int(0)), _dp(A), _dp(p), _dp(q), _dp(b), _dp(cd),
                                     meta.ctypes.data_as(ct.POINTER(ct.c_int)), cmap.ctypes.data_as(ct.POINTER(ct.c_uint8)), csmap.ctypes.data_as(ct.POINTER(ct.c_int8)),
                                     F.ctypes.data_as(ct.POINTER(ct.c_int)))
        return dict(A=A, p=p, q=q, b=b, cd=cd, meta=meta, cmap=cmap, csmap=csmap, F=F)


def _ssum(a):
    a = np.asarray(a, np.float64).ravel()
    return float(np.add.accumulate(a)[-1]) if a.size else 0.0


def run_scenario(lib, sc, full_upto=0, max_steps=None, capture=False, shift_mode="recorded", on_step=None, print_basic_info=False, split=0,
                 on_create=None, device=-1, lean=False, early_scale=0, fast_moments=False):
    """Returns {name: array} in the dump layout. capture=True adds the post-MUC term list / F arrays of full steps."""
    s = Session(lib, sc, print_basic_info=print_basic_info, split=split, device=device, lean=lean, early_scale=early_scale, fast_moments=fast_moments)
    if on_create is not None:
        on_create(s)
    out = {}
    d = sc.d
    MS = s.shape_range - 1
    try:
        if capture:
            lib.mce_debug_capture(s.h, 1)
        nrec = len(sc.rec) if max_steps is None else min(max_steps, len(sc.rec))
        for k in range(nrec):
            r = sc.rec[k]
            sp = "s%d" % (k + 1)
            first = k == 0
            with_tp = (k % sc.p) == 0 and not first
            err = s.step(r)
            mo = s.moments()
            st = s.stats()
            out[sp + "/info"] = np.array([int(with_tp), mo.skip_post_mu, mo.Nt_after_muc, mo.Nt, err, int(first)], np.int32)
            out[sp + "/muc/counts"] = s.counts(True)
            mom = np.zeros(1 + d + d * d, np.complex128)
            mom[0] = complex(mo.fz[0], mo.fz[1]) if first else complex(mo.fz_after_mu[0], mo.fz_after_mu[1])  # ref_run can only see fz = 1 after step 1
            mom[1 : 1 + d] = np.array(mo.mean[: 2 * d]).view(np.complex128)
            mom[1 + d :] = np.array(mo.cov[: 2 * d * d]).view(np.complex128)
            out[sp + "/moments"] = mom
            out[sp + "/gscale"] = np.array([mo.g_scale_factor])
            out[sp + "/stats"] = np.array([st.ms_total, st.ms_tp, st.ms_mu, st.ms_moments, st.ms_regroup, st.ms_ftr, st.ms_gtable, st.ms_compact,
                                           st.ftr_rounds_max, st.diag_unmodelled_alias, st.diag_hash_overflow, st.kernel_launches, st.split_groups])
            out[sp + "/lean/stats"] = np.array([st.gtable_lean_launches], np.int64)
            full = (k + 1) <= full_upto
            if not mo.skip_post_mu:
                cnt = s.counts(False)
                out[sp + "/ftr/counts"] = cnt
                for m in range(1, s.shape_range):
                    if cnt[m] <= 0:
                        continue
                    e = s.export_shape(m)
                    pre = "%s/ftr/m%d" % (sp, m)
                    out[pre + "/digest"] = key_digest(e["cells"], e["keys"])
                    out[pre + "/fdigest"] = np.array([_ssum(np.abs(e["G"])), _ssum(e["p"]), _ssum(np.abs(e["b"]))])   # serial sums, like ref_run.cpp
                    if full:
                        for nm in ("A", "p", "b", "cells", "keys", "G"):
                            out[pre + "/" + nm] = e[nm]
                        out[pre + "/encB"] = e["keys"].astype(np.int32)
                if full and capture and not first:
                    mc = s.counts(True)
                    for m in range(1, s.shape_range):
                        if mc[m] <= 0:
                            continue
                        c = s.debug_muc_shape(m)
                        if c is None:
                            continue
                        pre = "%s/muc/m%d" % (sp, m)
                        for nm in ("A", "p", "q", "b", "cd", "meta", "F"):
                            out[pre + "/" + nm] = c[nm]
                        cm = np.full((c["cmap"].shape[0], MS), 255, np.uint8)
                        cm[:, : min(MS, MAXM)] = c["cmap"][:, :MS]
                        out[pre + "/cmap"] = cm
                        out[pre + "/csmap"] = np.ascontiguousarray(c["csmap"][:, :MS])
            if on_step:
                on_step(k + 1, s, out)
            if r.shift_kind == SHIFT_EXPLICIT and shift_mode == "recorded":
                s.shift_b(r.delta, -1.0)
            elif r.shift_kind in (SHIFT_OWN_MEAN, SHIFT_EXPLICIT):
                s.shift_b(np.array(mo.mean[: 2 * d])[0::2], -1.0)
    finally:
        s.close()
    return out


def run_scenario_partitioned(lib, sc, dist, full_upto=0, max_steps=None, transport="callback", moments="ordered", device=-1, on_step=None):
    """ONE estimator partitioned over the ranks of `dist` (cauchyfriendly_b200/shard.py).  Returns the dump layout of
    run_scenario with the ranks' term lists merged into the canonical (one-GPU) order through mce_shard_export_gpos, so the
    result can be compared with the same golden dumps -- on every rank."""
    from cauchyfriendly_b200.shard import init_term_sharding
    s = Session(lib, sc, device=device)
    init_term_sharding(s.h, dist, lib=lib, transport=transport, device=device, moments=moments)
    out = {}
    d = sc.d
    world = dist.get_world_size()
    try:
        nrec = len(sc.rec) if max_steps is None else min(max_steps, len(sc.rec))
        for k in range(nrec):
            r = sc.rec[k]
            sp = "s%d" % (k + 1)
            first = k == 0
            with_tp = (k % sc.p) == 0 and not first
            err = s.step(r)
            mo = s.moments()
            out[sp + "/info"] = np.array([int(with_tp), mo.skip_post_mu, mo.Nt_after_muc, mo.Nt, err, int(first)], np.int32)
            out[sp + "/muc/counts"] = s.counts(True)
            mom = np.zeros(1 + d + d * d, np.complex128)
            mom[0] = complex(mo.fz[0], mo.fz[1]) if first else complex(mo.fz_after_mu[0], mo.fz_after_mu[1])
            mom[1 : 1 + d] = np.array(mo.mean[: 2 * d]).view(np.complex128)
            mom[1 + d :] = np.array(mo.cov[: 2 * d * d]).view(np.complex128)
            out[sp + "/moments"] = mom
            out[sp + "/gscale"] = np.array([mo.g_scale_factor])
            if not mo.skip_post_mu:
                cnt = s.counts(False)
                out[sp + "/ftr/counts"] = cnt
                nloc = lib.mce_shard_export_gpos(s.h, None, 0)
                gpos = np.zeros(max(nloc, 1), np.int32)
                lib.mce_shard_export_gpos(s.h, gpos.ctypes.data_as(ct.POINTER(ct.c_int)), nloc)
                gpos = gpos[:nloc]
                mine, off = {}, 0
                for m in range(1, s.shape_range):
                    if cnt[m] <= 0:
                        continue
                    e = s.export_shape(m)
                    n = e["A"].shape[0]
                    e["gpos"] = gpos[off:off + n].copy()
                    off += n
                    mine[m] = e
                assert off == nloc, (off, nloc)
                everyone = [None] * world
                dist.all_gather_object(everyone, mine)
                base = 0
                for m in range(1, s.shape_range):
                    if cnt[m] <= 0:
                        continue
                    parts = [ev[m] for ev in everyone if m in ev and ev[m]["A"].shape[0] > 0]
                    gp = np.concatenate([q["gpos"] for q in parts])
                    order = np.argsort(gp, kind="stable")
                    assert gp.size == cnt[m] and np.array_equal(gp[order], base + np.arange(cnt[m])), "global alive ranks of shape %d are not a permutation" % m
                    base += cnt[m]
                    cells = np.concatenate([q["cells"] for q in parts])
                    starts = np.concatenate([[0], np.cumsum(cells)])[:-1]
                    keys_l = np.concatenate([q["keys"] for q in parts]); G_l = np.concatenate([q["G"] for q in parts])
                    e = dict(A=np.concatenate([q["A"] for q in parts])[order], p=np.concatenate([q["p"] for q in parts])[order],
                             b=np.concatenate([q["b"] for q in parts])[order], cells=cells[order])
                    e["keys"] = np.concatenate([keys_l[starts[i]:starts[i] + cells[i]] for i in order]) if order.size else keys_l
                    e["G"] = np.concatenate([G_l[starts[i]:starts[i] + cells[i]] for i in order]) if order.size else G_l
                    pre = "%s/ftr/m%d" % (sp, m)
                    out[pre + "/digest"] = key_digest(e["cells"], e["keys"])
                    out[pre + "/fdigest"] = np.array([_ssum(np.abs(e["G"])), _ssum(e["p"]), _ssum(np.abs(e["b"]))])
                    if (k + 1) <= full_upto:
                        for nm in ("A", "p", "b", "cells", "keys", "G"):
                            out[pre + "/" + nm] = e[nm]
                        out[pre + "/encB"] = e["keys"].astype(np.int32)
            if on_step:
                on_step(k + 1, s, out)
            if r.shift_kind == SHIFT_EXPLICIT:
                s.shift_b(r.delta, -1.0)
            elif r.shift_kind == SHIFT_OWN_MEAN:
                s.shift_b(np.array(mo.mean[: 2 * d])[0::2], -1.0)
    finally:
        s.close()
    return out


def oracle_dump(scenario_path, out_path, full_upto=0, max_steps=None, use_ref=False, print_basic_info=False):
    """Runs the C oracle (or the compiled reference) on a scenario file and reads its dump."""
    exe = os.path.join(ROOT, "oracle", "_ref", "ref_run_cpu1") if use_ref else os.path.join(ROOT, "oracle", "_build", "mce_oracle_run")
    cmd = [exe, scenario_path, out_path, "--full-upto", str(full_upto)]
    if max_steps is not None:
        cmd += ["--max-steps", str(max_steps)]
    if print_basic_info:
        cmd += ["--print-basic-info"]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL)
    return read_dump(out_path)


def cpdf_steps(gold):
    """Steps for which a golden / oracle cpdf dump holds `s<k>/cpdf1d/i<idx>` arrays."""
    return sorted({int(n.split("/")[0][1:]) for n in gold if "/cpdf1d/i" in n})


def run_cpdf1d(lib, sc, gold, max_step=None):
    """Replays the scenario and evaluates the point-wise 1-D marginal cpdf of every state after each step the golden dump
    covers (after the recorded shift, like oracle/ref_cpdf.cpp), on the golden dump's grid with its bar_nu.
    Returns {name: [n][2]} in the dump layout, through mce_marginal_1d_grid of the C ABI."""
    lo, hi, res = [float(v) for v in gold["cpdf1d/grid"]]
    bar_nu = np.ascontiguousarray(gold["cpdf1d/bar_nu"], np.float64)
    want = [k for k in cpdf_steps(gold) if max_step is None or k <= max_step]
    s = Session(lib, sc)
    out = {}
    try:
        for k in range(max(want)):
            r = sc.rec[k]
            s.step(r)
            mo = s.moments()
            if r.shift_kind == SHIFT_EXPLICIT:
                s.shift_b(r.delta, -1.0)
            elif r.shift_kind == SHIFT_OWN_MEAN:
                s.shift_b(np.array(mo.mean[: 2 * sc.d])[0::2], -1.0)
            if (k + 1) not in want:
                continue
            n = lib.mce_cpdf_grid_count(lo, hi, res)
            for idx in range(sc.d):
                xy = np.zeros((n, 2))
                rc = lib.mce_marginal_1d_grid(s.h, idx, _dp(bar_nu), lo, hi, res, _dp(xy), n)
                if rc != n:
                    raise RuntimeError("mce_marginal_1d_grid returned %d: %s" % (rc, lib.mce_last_error().decode()))
                out["s%d/cpdf1d/i%d" % (k + 1, idx)] = xy
    finally:
        s.close()
    return out


def oracle_cpdf1d(scenario_path, out_path, lo, hi, res, steps, use_ref=False):
    """1-D marginal cpdf grids from the C oracle (or the compiled reference, oracle/_ref/ref_cpdf_cpu1)."""
    st = ",".join(str(k) for k in steps)
    if use_ref:
        cmd = [os.path.join(ROOT, "oracle", "_ref", "ref_cpdf_cpu1"), scenario_path, out_path, repr(lo), repr(hi), repr(res), st]
    else:
        cmd = [os.path.join(ROOT, "oracle", "_build", "mce_oracle_run"), scenario_path, out_path, "--max-steps", str(max(steps)),
               "--cpdf1d", repr(lo), repr(hi), repr(res), st]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return {n: v for n, v in read_dump(out_path).items() if "cpdf1d" in n}


def run_cpdf2d(lib, sc, gold, max_step=None):
    """2-D marginal grids of every state pair the golden dump holds (`s<k>/cpdf2d/i<a>_<b>`), through mce_marginal_2d_grid."""
    xlo, xhi, xres, ylo, yhi, yres = [float(v) for v in gold["cpdf2d/grid"]]
    bar_nu = np.ascontiguousarray(gold["cpdf1d/bar_nu"], np.float64)
    names = [n for n in gold if "/cpdf2d/i" in n and (max_step is None or int(n.split("/")[0][1:]) <= max_step)]
    want = sorted({int(n.split("/")[0][1:]) for n in names})
    s = Session(lib, sc)
    out = {}
    try:
        for k in range(max(want)):
            r = sc.rec[k]
            s.step(r)
            mo = s.moments()
            if r.shift_kind == SHIFT_EXPLICIT:
                s.shift_b(r.delta, -1.0)
            elif r.shift_kind == SHIFT_OWN_MEAN:
                s.shift_b(np.array(mo.mean[: 2 * sc.d])[0::2], -1.0)
            for nme in [n for n in names if int(n.split("/")[0][1:]) == k + 1]:
                a, b = [int(v) for v in nme.split("/i")[1].split("_")]
                n = lib.mce_cpdf_grid_count(xlo, xhi, xres) * lib.mce_cpdf_grid_count(ylo, yhi, yres)
                xyz = np.zeros((n, 3))
                nx, ny = ct.c_int(0), ct.c_int(0)
                rc = lib.mce_marginal_2d_grid(s.h, a, b, _dp(bar_nu), xlo, xhi, xres, ylo, yhi, yres, _dp(xyz), n, ct.byref(nx), ct.byref(ny))
                if rc != n:
                    raise RuntimeError("mce_marginal_2d_grid returned %d: %s" % (rc, lib.mce_last_error().decode()))
                out[nme] = xyz
    finally:
        s.close()
    return out


def run_transforms(lib, sc, gold, k0):
    """deterministic_time_prop / shift_cf_by_bias after k0 steps (oracle/ref_transforms.cpp): returns (arrays, problems) where
    arrays follow the golden dump's names (t1/t2/t3 term lists, later moments) through the C ABI."""
    d = sc.d
    T = np.ascontiguousarray(gold["T"], np.float64)
    bias = np.ascontiguousarray(gold["bias"], np.float64)
    s = Session(lib, sc)
    out = {}

    def do_step(k):
        r = sc.rec[k]
        s.step(r)
        mo = s.moments()
        if r.shift_kind == SHIFT_EXPLICIT:
            s.shift_b(r.delta, -1.0)
        elif r.shift_kind == SHIFT_OWN_MEAN:
            s.shift_b(np.array(mo.mean[: 2 * d])[0::2], -1.0)
        return mo

    def dump(pre):
        cnt = s.counts(False)
        for m in range(1, s.shape_range):
            if cnt[m] > 0:
                e = s.export_shape(m)
                for nm in ("A", "p", "b"):
                    out["%s/m%d/%s" % (pre, m, nm)] = e[nm]

    try:
        for k in range(k0):
            do_step(k)
        assert lib.mce_deterministic_time_prop(s.h, _dp(T), None, None) == 0
        dump("t1")
        assert lib.mce_shift_b(s.h, _dp(bias), 1.0) == 0            # shift_cf_by_bias, est:1312
        dump("t2")
        r0 = sc.rec[k0]
        if r0.B is not None:
            B = np.ascontiguousarray(r0.B, np.float64); u = np.ascontiguousarray(r0.u, np.float64)
            assert lib.mce_deterministic_time_prop(s.h, _dp(T), _dp(B), _dp(u)) == 0
            dump("t3")
        for k in range(k0, len(sc.rec)):
            mo = do_step(k)
            mom = np.zeros(1 + d + d * d, np.complex128)
            mom[0] = complex(mo.fz[0], mo.fz[1])                         # the public field est.fz after the whole step
            mom[1:1 + d] = np.array(mo.mean[: 2 * d]).view(np.complex128)
            mom[1 + d:] = np.array(mo.cov[: 2 * d * d]).view(np.complex128)
            out["s%d/moments" % (k + 1)] = mom
            out["s%d/info" % (k + 1)] = np.array([mo.Nt, mo.numeric_moment_errors], np.int32)
    finally:
        s.close()
    return out

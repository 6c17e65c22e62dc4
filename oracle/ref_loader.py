# -*- coding: utf-8 -*-
"""Import the UNMODIFIED reference from /root/reference  --  TEST INFRASTRUCTURE ONLY.

Used by ``tests/golden/make_golden.py`` (fixture generation) and by the
in-container oracle tests.  The reference tree does not exist on the GPU box;
everything here raises ``ReferenceUnavailable`` there and the callers skip.

The reference needs three packages that are not installed in this image; they
are replaced by minimal stand-ins injected through ``sys.modules``:

* ``safer``      only used by ``common.dump_output`` (common.py:150-152) -> empty module
* ``autograd``   ``autograd.numpy`` -> numpy; ``jacobian`` -> complex-step
                 (exact to rounding for the analytic nonlinearities used);
                 ``grad`` -> central differences (theta gradient only)
* ``matplotlib`` only used by ``psmf/tracking.py`` for figures -> inert stub
"""

from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np

REF_ROOT = os.environ.get("RPSMF_REF", "/root/reference")


class ReferenceUnavailable(RuntimeError):
    pass


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ExperimentImpute"))


def _jacobian(fun, argnum=0):
    def jac(*args):
        x = np.asarray(args[argnum], dtype=np.float64)
        out0 = np.asarray(fun(*args))
        J = np.zeros(out0.shape + x.shape)
        h = 1e-30
        it = np.nditer(x, flags=["multi_index"])
        for _ in it:
            xc = x.astype(np.complex128)
            xc[it.multi_index] += 1j * h
            a = list(args)
            a[argnum] = xc
            J[(Ellipsis,) + it.multi_index] = np.imag(np.asarray(fun(*a))) / h
        return J
    return jac


def _grad(fun, argnum=0):
    def g(*args):
        x = np.asarray(args[argnum], dtype=np.float64)
        out = np.zeros(x.shape)
        it = np.nditer(x, flags=["multi_index"])
        for _ in it:
            h = 1e-6 * max(1.0, abs(float(x[it.multi_index])))
            ap = list(args); am = list(args)
            xp = x.copy(); xm = x.copy()
            xp[it.multi_index] += h; xm[it.multi_index] -= h
            ap[argnum] = xp; am[argnum] = xm
            out[it.multi_index] = (np.asarray(fun(*ap)).item() - np.asarray(fun(*am)).item()) / (2 * h)
        return out
    return g


def install_stubs():
    if "safer" not in sys.modules:
        sys.modules["safer"] = types.ModuleType("safer")
    if "autograd" not in sys.modules:
        ag = types.ModuleType("autograd")
        ag.numpy = np
        ag.grad = _grad
        ag.jacobian = _jacobian
        sys.modules["autograd"] = ag
        sys.modules["autograd.numpy"] = np
    try:
        import matplotlib  # noqa: F401
    except Exception:
        class _Any:
            def __getattr__(self, k):
                return _Any()

            def __call__(self, *a, **k):
                return _Any()

            def __setitem__(self, k, v):
                pass

            def __getitem__(self, k):
                return _Any()

            def __iter__(self):
                return iter(())

        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.rcParams = {}
        plt.__getattr__ = lambda name: _Any()   # type: ignore[attr-defined]
        mpl.pyplot = plt
        mpl.__getattr__ = lambda name: _Any()   # type: ignore[attr-defined]
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt


def _load(path, name, extra_path):
    if not available():
        raise ReferenceUnavailable(REF_ROOT)
    install_stubs()
    for p in reversed(extra_path):
        if p not in sys.path:
            sys.path.insert(0, p)
    # joblib.Memory("./cache") is created at import (rPSMF.py:27): import from a scratch cwd
    cwd = os.getcwd()
    scratch = tempfile.mkdtemp(prefix="rpsmf_ref_")
    os.chdir(scratch)
    try:
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        sys.modules[name] = mod
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def impute_module(which: str):
    """which in {'rPSMF', 'PSMF', 'common'} -> module object of ExperimentImpute/<which>.py"""
    base = os.path.join(REF_ROOT, "ExperimentImpute")
    return _load(os.path.join(base, which + ".py"), "ref_impute_" + which, [base])


def pypsmf():
    """The reference ``psmf`` package (pypsmf/psmf)."""
    if not available():
        raise ReferenceUnavailable(REF_ROOT)
    install_stubs()
    p = os.path.join(REF_ROOT, "pypsmf")
    if p not in sys.path:
        sys.path.insert(0, p)
    return importlib.import_module("psmf")


def synthetic_module(which: str):
    """which in {'synthetic_psmf', 'synthetic_rpsmf', 'data'} (ExperimentSynthetic)."""
    pypsmf()
    base = os.path.join(REF_ROOT, "ExperimentSynthetic")
    return _load(os.path.join(base, which + ".py"), "ref_syn_" + which, [base])


class RecordingNumpy:
    """Drop-in for the ``np`` global of a reference module that records the
    arguments of a few calls, so per-step internals of the unmodified loop body
    (rPSMF.py:81-135 / PSMF.py:60-84) can be dumped without editing it:

    * ``np.linalg.inv``  1st call per step: PP = P + Q (rPSMF.py:34,87);
                         2nd: Pi + CM' Ri CM (rPSMF.py:35-36)
    * ``np.trace``       eta_k * d (rPSMF.py:108 / PSMF.py:77)
    * ``np.sqrt``        diag(U) (rPSMF.py:121) or Nt (PSMF.py:83-84)
    """

    def __init__(self):
        self.inv_args = []
        self.trace_vals = []
        self.sqrt_args = []
        outer = self

        class _Linalg:
            def __getattr__(self, k):
                return getattr(np.linalg, k)

            def inv(self, A):
                outer.inv_args.append(np.array(A, copy=True))
                return np.linalg.inv(A)

        self.linalg = _Linalg()

    def __getattr__(self, k):
        return getattr(np, k)

    def trace(self, A):
        v = np.trace(A)
        self.trace_vals.append(float(v))
        return v

    def sqrt(self, A):
        self.sqrt_args.append(np.array(A, copy=True))
        return np.sqrt(A)

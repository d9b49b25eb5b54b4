"""Minimal stand-in for the `lmfit` package (TEST INFRASTRUCTURE, not product code).

The reference imports lmfit at /root/reference/python/simpleicp/optimization.py:12 and uses
exactly this subset (call sites optimization.py:72, :80-90, :93-101, :104-111, :117-119,
:139-159, :165-170):

    lmfit.Parameters()                      ordered mapping name -> Parameter
    lmfit.Parameter(name=, value=, vary=, user_data=)
    lmfit.minimize(fcn, params, method="least_squares", args=(...))
        -> result.params[name].value / .vary / .user_data, result.residual, result.jac

lmfit is an unpinned dependency of the reference (python/setup.py:24) that is not installed in
this image and cannot be installed (no network).  Upstream lmfit's Minimizer.least_squares
deep-copies the parameters, optimises the vary=True ones in insertion order through
scipy.optimize.least_squares(residual, x0, bounds=(-inf, inf), max_nfev=2*2000*(nvarys+1))
with SciPy's defaults (method='trf', jac='2-point', ftol=xtol=gtol=1e-8, x_scale=1,
loss='linear'), re-evaluates the residual at the solution and exposes ret.jac.  That published
behaviour is what is restated here.
"""
import copy

import numpy as np
from scipy.optimize import least_squares

__version__ = "standin-0"


class Parameter:
    def __init__(self, name=None, value=None, vary=True, user_data=None, **_ignored):
        self.name = name
        self.value = float(value) if value is not None else None
        self.vary = bool(vary)
        self.user_data = user_data
        self.stderr = None


class Parameters(dict):
    """Insertion-ordered name -> Parameter mapping."""


class MinimizerResult:
    pass


def minimize(fcn, params, method="leastsq", args=None, kws=None, **fit_kws):
    if method != "least_squares":
        raise NotImplementedError("stand-in implements method='least_squares' only")
    args = tuple(args or ())
    kws = dict(kws or {})
    work = copy.deepcopy(params)
    names = [n for n in work if work[n].vary]
    x0 = np.array([work[n].value for n in names], dtype=float)

    def residual(x):
        for n, v in zip(names, x):
            work[n].value = float(v)
        return np.asarray(fcn(work, *args, **kws), dtype=float)

    res = MinimizerResult()
    if len(names) == 0:
        res.params = work
        res.residual = residual(x0)
        res.jac = np.zeros((res.residual.size, 0))
        res.nfev = 1
        return res
    ret = least_squares(
        residual, x0, bounds=(-np.inf, np.inf), max_nfev=2 * 2000 * (len(names) + 1), **fit_kws
    )
    res.residual = residual(ret.x)  # lmfit re-evaluates at the solution
    res.params = work
    res.jac = ret.jac
    res.nfev = ret.nfev
    res.success = ret.success
    res.message = ret.message
    return res

"""lmfit compatibility layer.

ExTrack's public API takes and returns lmfit objects (``Parameters`` in,
``MinimizerResult`` out; reference ``extrack/tracking.py:31,1210-1211,1287-1288,1371``).
When lmfit is importable it is used unchanged.  This image has no lmfit (and no
network), so a small stand-in with the same surface is provided:

* ``Parameters`` / ``Parameter`` with ``value, min, max, vary, expr, brute_step``;
  ``expr`` constraints are re-evaluated whenever a value is read, iteration order
  is insertion order (the reference relies on both, ``tracking.py:944,1073``).
* ``minimize(fcn, params, args, method, nan_policy)`` for a scalar objective with
  lmfit's bound transform (two-sided: ``arcsin``; one-sided: ``sqrt``) on top of
  ``scipy.optimize.minimize`` ('bfgs', 'powell', 'nelder', 'l-bfgs-b', 'cg').

Fitted-parameter parity against real lmfit is unpinned in this image (SURVEY.md
§0 hazard #3); the stand-in follows lmfit's documented MINPACK-style transform.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - exercised only where lmfit exists
    from lmfit import Parameters, Parameter, minimize  # type: ignore # noqa: F401

    HAVE_LMFIT = True
except Exception:  # lmfit absent -> stand-in below
    HAVE_LMFIT = False


if not HAVE_LMFIT:

    _SAFE_FUNCS = {
        k: getattr(math, k)
        for k in ("sqrt", "exp", "log", "log10", "sin", "cos", "tan", "asin", "acos", "atan", "pi", "e")
    }
    _SAFE_FUNCS.update({"abs": abs, "min": min, "max": max})

    _EXPR_CODE = {}

    class Parameter:
        """One named fit parameter (subset of ``lmfit.Parameter``)."""

        def __init__(self, name, value=None, vary=True, min=-np.inf, max=np.inf, expr=None, brute_step=None):
            self.name = name
            self._val = value
            self.vary = bool(vary) if expr is None else False
            self.min = -np.inf if min is None else min
            self.max = np.inf if max is None else max
            self.expr = expr
            self.brute_step = brute_step
            self.stderr = None
            self._owner = None
            self.init_value = value
            if self._val is not None and expr is None:
                self._val = float(np.clip(self._val, self.min, self.max))

        # value resolves expr constraints lazily against the owning Parameters
        @property
        def value(self):
            if self.expr is not None and self._owner is not None:
                return self._owner._eval(self.expr)
            return self._val

        @value.setter
        def value(self, v):
            self._val = v

        def set(self, value=None, vary=None, min=None, max=None, expr=None, brute_step=None):
            if value is not None:
                self._val = value
            if vary is not None:
                self.vary = vary
            if min is not None:
                self.min = min
            if max is not None:
                self.max = max
            if expr is not None:
                self.expr = expr
                self.vary = False
            if brute_step is not None:
                self.brute_step = brute_step

        def __float__(self):
            return float(self.value)

        def __repr__(self):
            return f"<Parameter '{self.name}', value={self.value}, bounds=[{self.min}:{self.max}], vary={self.vary}, expr={self.expr!r}>"

    class Parameters(OrderedDict):
        """Ordered name -> Parameter mapping (subset of ``lmfit.Parameters``)."""

        def add(self, name, value=None, vary=True, min=-np.inf, max=np.inf, expr=None, brute_step=None):
            if isinstance(name, Parameter):
                p = name
            else:
                p = Parameter(name, value=value, vary=vary, min=min, max=max, expr=expr, brute_step=brute_step)
            p._owner = self
            OrderedDict.__setitem__(self, p.name, p)
            return p

        def add_many(self, *parlist):
            for spec in parlist:
                self.add(*spec)

        def _eval(self, expr, _depth=0):
            if _depth > 32:
                raise RecursionError("circular parameter expression: %r" % expr)
            names = dict(_SAFE_FUNCS)
            for k, p in OrderedDict.items(self):
                if p.expr is None:
                    names[k] = p._val
            # resolve nested expressions on demand
            code = _EXPR_CODE.get(expr)
            if code is None:
                code = _EXPR_CODE[expr] = compile(expr, "<expr>", "eval")  # (constraint expressions are evaluated at every objective call)
            for n in code.co_names:
                if n not in names and n in self:
                    names[n] = self._eval(OrderedDict.__getitem__(self, n).expr, _depth + 1)
            return eval(code, {"__builtins__": {}}, names)

        def valuesdict(self):
            return OrderedDict((k, p.value) for k, p in OrderedDict.items(self))

        def copy(self):
            return self.__deepcopy__({})

        def __deepcopy__(self, memo):
            out = Parameters()
            for k, p in OrderedDict.items(self):
                q = Parameter(k, value=p._val, vary=p.vary, min=p.min, max=p.max, expr=p.expr, brute_step=p.brute_step)
                q.vary = p.vary
                q._val = p._val
                q.stderr = p.stderr
                q.init_value = p.init_value
                out.add(q)
            return out

        def pretty_print(self):  # minimal
            for k, p in OrderedDict.items(self):
                print(f"{k:>16s} {p.value!r:>24} min={p.min} max={p.max} vary={p.vary} expr={p.expr}")

    # ---- bound transforms (lmfit / MINPACK-1 convention) -------------------------------
    def _to_internal(p):
        v, lo, hi = p._val, p.min, p.max
        if np.isfinite(lo) and np.isfinite(hi):
            return math.asin(2.0 * (v - lo) / (hi - lo) - 1.0)
        if np.isfinite(lo):
            return math.sqrt((v - lo + 1.0) ** 2 - 1.0)
        if np.isfinite(hi):
            return math.sqrt((hi - v + 1.0) ** 2 - 1.0)
        return v

    def _from_internal(p, x):
        lo, hi = p.min, p.max
        if np.isfinite(lo) and np.isfinite(hi):
            return lo + (math.sin(x) + 1.0) * (hi - lo) / 2.0
        if np.isfinite(lo):
            return lo - 1.0 + math.sqrt(x * x + 1.0)
        if np.isfinite(hi):
            return hi + 1.0 - math.sqrt(x * x + 1.0)
        return x

    class MinimizerResult:
        """Result container (subset of ``lmfit.minimizer.MinimizerResult``)."""

        def __init__(self):
            self.params = None
            self.residual = None
            self.success = False
            self.message = ""
            self.nfev = 0
            self.method = ""
            self.init_vals = []
            self.var_names = []

    _METHODS = {
        "bfgs": "BFGS",
        "bfsg": "BFGS",  # the tutorial's misspelling is tolerated (Tutorial_ExTrack.ipynb cell 50)
        "powell": "Powell",
        "nelder": "Nelder-Mead",
        "nelder-mead": "Nelder-Mead",
        "lbfgsb": "L-BFGS-B",
        "l-bfgs-b": "L-BFGS-B",
        "cg": "CG",
        "cobyla": "COBYLA",
        "tnc": "TNC",
        "slsqp": "SLSQP",
    }

    def minimize(fcn, params, method="leastsq", args=None, kws=None, nan_policy="raise", max_nfev=None, **fit_kws):
        """Scalar-objective minimisation with lmfit semantics (see module docstring)."""
        from scipy.optimize import minimize as _spmin

        args = tuple(args or ())
        kws = dict(kws or {})
        work = params.copy()
        free = [p for p in work.values() if p.vary and p.expr is None]
        res = MinimizerResult()
        res.var_names = [p.name for p in free]
        res.init_vals = [p._val for p in free]
        res.method = method
        x0 = np.array([_to_internal(p) for p in free], dtype=float)
        count = [0]

        def penalty(x):
            for p, xi in zip(free, x):
                p._val = _from_internal(p, float(xi))
            count[0] += 1
            r = fcn(work, *args, **kws)
            if isinstance(r, np.ndarray) and r.size > 1:
                r = float((r * r).sum())
            r = float(np.asarray(r).reshape(-1)[0]) if not np.isscalar(r) else float(r)
            if nan_policy == "raise" and not np.isfinite(r):
                raise ValueError("objective returned a non-finite value (nan_policy='raise')")
            return r

        spm = _METHODS.get(str(method).lower(), method)
        opts = dict(fit_kws.pop("options", {}))
        if max_nfev is not None:
            opts.setdefault("maxiter", int(max_nfev))
        if len(free) == 0:
            fbest = penalty(x0)
            res.success, res.message = True, "no free parameters"
        else:
            out = _spmin(penalty, x0, method=spm, options=opts, **fit_kws)
            for p, xi in zip(free, np.atleast_1d(out.x)):
                p._val = _from_internal(p, float(xi))
            fbest = float(out.fun)
            res.success, res.message = bool(out.success), str(out.message)
        res.nfev = count[0]
        res.params = work
        res.residual = np.atleast_1d(np.asarray(fbest, dtype=float))
        return res

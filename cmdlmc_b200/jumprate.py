"""Jump-rate functions -- mirror of mdlmc/LMC/jumprate_generators.py:14-43 plus the legacy
activation-energy / exponential kinds described in mdlmc/IO/config_parser.py:322-349.

Calling an instance evaluates the rate on the GPU through the C ABI; `kind` / `params` let the
fused topology kernels evaluate the same function without leaving the device."""
from abc import ABCMeta

import numpy as np

from . import _abi, runtime
from ._abi import as_f64, check, ptr

NPAR = 8


class JumpRate(metaclass=ABCMeta):
    """Calculates a proton hopping rate as a function of geometric parameters such as distance,
    angle, etc."""
    kind = None

    @property
    def params(self):
        raise NotImplementedError

    def _par(self):
        p = list(self.params)
        return as_f64(p + [0.0] * (NPAR - len(p)))

    def _eval(self, x, theta=None):
        runtime.ensure_init()
        x = np.asarray(x, dtype=float)
        flat = as_f64(x.reshape(-1))
        th = as_f64(np.asarray(theta, dtype=float).reshape(-1)) if theta is not None else None
        out = np.empty_like(flat)
        check(_abi.lib().cmd_rates(self.kind, ptr(self._par()), ptr(flat),
                                   ptr(th) if th is not None else None, flat.shape[0], ptr(out)))
        return out.reshape(x.shape)


class Fermi(JumpRate):
    __show_in_config__ = True
    kind = 0

    def __init__(self, a: float, b: float, c: float):
        """a: Amplitude, b: Location, c: Width"""
        self._a = a
        self._b = b
        self._c = c

    @property
    def params(self):
        return (self._a, self._b, self._c)

    def __call__(self, x):
        return self._eval(x)


class FermiAngle(Fermi):
    kind = 1

    def __init__(self, a: float, b: float, c: float, theta: float):
        super().__init__(a, b, c)
        self._theta = theta

    @property
    def params(self):
        return (self._a, self._b, self._c, self._theta)

    def __call__(self, x, theta):
        return self._eval(x, theta)


class ActivationEnergy(JumpRate):
    """Legacy "AE_rates" (IO/config_parser.py:334-342; parity unpinned):
    E(d) = a (d - d0) / sqrt(b + 1 / (d - d0)^2),  w(d) = A exp(-E(d) / (k_B T)),  w = A for
    d <= d0;  k_B = 8.617333262e-5 eV/K."""
    __show_in_config__ = True
    kind = 2

    def __init__(self, A: float, a: float, b: float, d0: float, T: float):
        self._A, self._a, self._b, self._d0, self._T = A, a, b, d0, T

    @property
    def params(self):
        return (self._A, self._a, self._b, self._d0, self._T)

    def __call__(self, x):
        return self._eval(x)


class Exponential(JumpRate):
    """Legacy "Exponential_rates" (IO/config_parser.py:344-345): w(d) = a exp(b d)."""
    __show_in_config__ = True
    kind = 3

    def __init__(self, a: float, b: float):
        self._a, self._b = a, b

    @property
    def params(self):
        return (self._a, self._b)

    def __call__(self, x):
        return self._eval(x)

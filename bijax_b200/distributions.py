"""Declarative stand-ins for the tfp.distributions that bijax priors are written with (README.md:28-52).  They only carry
parameters; the arithmetic (log_prob of the transformed prior and its gradient) runs in the CUDA prologue kernel.  The
engine duck-types on the class name and the TFP attribute names, so real tfd objects are accepted as well.
"""
from __future__ import annotations


class Distribution:
    pass


class Normal(Distribution):
    def __init__(self, loc, scale):
        self.loc, self.scale = loc, scale


class MultivariateNormalDiag(Distribution):
    def __init__(self, loc=None, scale_diag=None):
        self.loc, self.scale_diag = loc, scale_diag


class Beta(Distribution):
    def __init__(self, concentration1, concentration0):
        self.concentration1, self.concentration0 = concentration1, concentration0


class Gamma(Distribution):
    def __init__(self, concentration, rate):
        self.concentration, self.rate = concentration, rate

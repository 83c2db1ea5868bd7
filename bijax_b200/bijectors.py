"""Declarative stand-ins for the tfp.bijectors that bijax's ``bijector`` / ``prior_constraints`` dicts hold
(README.md:58-77; core.py:107-111,154-155).  TFP is not installable here; the engine recognises bijectors by duck typing
(class name, ``.bijectors`` for chains), so real tfb objects work unchanged where TFP exists.
"""
from __future__ import annotations

from typing import Sequence


class Bijector:
    pass


class Identity(Bijector):
    pass


class Exp(Bijector):
    pass


class Softplus(Bijector):
    pass


class Sigmoid(Bijector):
    pass


class Chain(Bijector):
    """tfb.Chain([b1, b2, ...]) = b1 o b2 o ...: the LAST bijector in the list is applied first."""

    def __init__(self, bijectors: Sequence[Bijector]):
        self.bijectors = list(bijectors)

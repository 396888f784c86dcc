"""Averaging probabilities over the k shadowing paths.

The reference imports `Softmax`, `Uniform`, `DiscreteProba` from the un-vendored
`scatspectra` v2.0.2 (path_shadowing.py:9); their source is not in the reference tree, so the
weight formula here is PARITY-UNPINNED: w ∝ exp(-d² / (2 η²)) -- "the width of a Gaussian in
the Gaussian average" (plot_utils.py:59-65).  It is isolated in `softmax_weights`.
Call-site contracts honoured (path_shadowing.py:228-230,251-252; plot_utils.py:74-76):
`Softmax(distances, eta)`, `Uniform()`, `.avg(x, axis)`, `.std(x, axis)`; weights of shape
(B,k,1) against x (B,k,nT) over axis 1, and (k,) against (k,1,T) over axis 0.
"""
from __future__ import annotations

import numpy as np


def softmax_weights(distances: np.ndarray, eta: float) -> np.ndarray:
    """Un-normalised Gaussian weights, shifted by the smallest d² (cancels on normalisation)."""
    d2 = np.square(np.asarray(distances, dtype=np.float64))
    return np.exp(-(d2 - d2.min()) / (2.0 * float(eta) ** 2))


class DiscreteProba:
    """A probability on the path axis given by non-negative weights (None = uniform)."""

    def __init__(self, weights: np.ndarray | None = None):
        self.weights = weights

    def _w(self, x: np.ndarray, axis: int) -> np.ndarray:
        axis = axis % x.ndim
        if self.weights is None:
            shape = [1] * x.ndim
            shape[axis] = x.shape[axis]
            return np.full(shape, 1.0 / x.shape[axis])
        w = np.asarray(self.weights, dtype=np.float64)
        if w.ndim < x.ndim:  # align the leading axes, pad trailing singleton axes
            lead = axis - (w.ndim - 1) if w.ndim - 1 <= axis else 0
            w = w.reshape((1,) * lead + w.shape + (1,) * (x.ndim - w.ndim - lead))
        return w / w.sum(axis=axis, keepdims=True)

    def avg(self, x: np.ndarray, axis: int = 0) -> np.ndarray:
        x = np.asarray(x)
        out = (self._w(x, axis) * x).sum(axis)
        return out.astype(x.dtype) if np.issubdtype(x.dtype, np.floating) else out

    def std(self, x: np.ndarray, axis: int = 0) -> np.ndarray:
        x = np.asarray(x)
        w = self._w(x, axis)
        xd = x.astype(np.float64)
        m = (w * xd).sum(axis)
        var = (w * xd * xd).sum(axis) - m * m
        out = np.sqrt(np.maximum(var, 0.0))
        return out.astype(x.dtype) if np.issubdtype(x.dtype, np.floating) else out


class Uniform(DiscreteProba):

    def __init__(self):
        super().__init__(None)


class Softmax(DiscreteProba):

    def __init__(self, distances: np.ndarray, eta: float):
        self.distances = np.asarray(distances)
        self.eta = eta
        super().__init__(softmax_weights(self.distances, eta) if self.distances.ndim <= 1
                         else self._per_query(self.distances, eta))

    @staticmethod
    def _per_query(d: np.ndarray, eta: float) -> np.ndarray:
        # (B, k, ...) : independent weights per query along axis 1
        d2 = np.square(d.astype(np.float64))
        return np.exp(-(d2 - d2.min(axis=1, keepdims=True)) / (2.0 * float(eta) ** 2))

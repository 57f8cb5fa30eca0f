"""Shim: type names referenced in annotations at tapnet/torch/utils.py:239-250."""


class Reshape:  # pragma: no cover - annotation only
  shape = ()


class Transpose:  # pragma: no cover - annotation only
  perm = ()


class Broadcast:  # pragma: no cover - annotation only
  axis_sizes = {}

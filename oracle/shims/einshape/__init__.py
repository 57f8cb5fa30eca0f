"""Import shim (test infrastructure): lets the unmodified reference torch path import.

The reference (`tapnet/torch/utils.py:19-20`) imports `einshape.src.{abstract_ops,backend}`;
the package is not installed in this image. Only `Backend.exec(equation, value, shape,
**sizes)` is used (`tapnet/torch/utils.py:272`), which is a pure layout operation, so an
einops-backed stand-in is sufficient. No reference code is copied here.
"""

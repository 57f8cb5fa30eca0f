"""Shim for `einshape.src.backend.Backend` used at tapnet/torch/utils.py:234-272.

`exec(equation, value, shape, **sizes)` re-groups / re-orders axes. We translate the
einshape equation (single-letter indices, e.g. 'tbnhw->(tbn)hw1', 'bn...->(bn)...') to
an einops pattern by putting spaces between the letters and call einops.rearrange.
"""
from typing import Generic, TypeVar

import einops

T = TypeVar('T')


def _space(side: str) -> str:
  out = []
  i = 0
  while i < len(side):
    if side.startswith('...', i):
      out.append('...')
      i += 3
    else:
      out.append(side[i])
      i += 1
  s = ' '.join(out)
  return s.replace('( ', '(').replace(' )', ')')


class Backend(Generic[T]):

  def exec(self, equation, value, shape, **sizes):  # noqa: A003 - name fixed by caller
    del shape
    lhs, rhs = equation.split('->')
    return einops.rearrange(value, f'{_space(lhs)} -> {_space(rhs)}', **sizes)

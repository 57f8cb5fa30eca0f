"""Shim for dm-tree's `map_structure` (tapnet/torch/tapir_model.py:27,786-802)."""
from torch.utils._pytree import tree_map


def map_structure(fn, *structures):
  return tree_map(fn, *structures)

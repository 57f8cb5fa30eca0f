"""JAX / Haiku `.npy` TAPIR checkpoints -> torch state dict.

"Online TAPIR" and the original TAPIR are published as pickled Haiku parameter trees
(`{'params': {module_name: {param_name: array}}, 'state': ...}`, README.md:161-171; loaded by
`np.load(path, allow_pickle=True).item()` in tapir_clustering.py:923-924) while the hot path
(tapnet/torch/tapir_model.py and this package) takes the torch layout (`tapnet_b200/schema.py`).
The reference ships no converter for TAPIR (only TAPNext's, tapnext_torch_utils.py:60).

Module names follow Haiku's rules applied to the JAX model definitions:
  models/tapir_model.py:318-390   modules built in TAPIR.__init__  -> 'tapir/~/<name>'
  models/resnet.py:160-262,385-450 ResNet / BlockGroup / BlockV2 build children in __init__
                                   -> 'tapir/~/resnet/~/block_group_g/~/block_b/~/conv_0'
  models/tapir_model.py:40-160    mixer children are built in __call__ with default names
                                   -> 'tapir/~/pips_mlp_mixer/block_3/mlp1_up_1', 'linear_1', ...
  models/tapir_model.py:162-186   ExtraConvs: 'layer_norm[_i]', 'conv2_d[_j]' (j = 2i, 2i+1)
(the mixer prefix is confirmed by the causal-state keys in tapir_clustering.py:824-847).  Keys are
matched after dropping the '~' path elements, so a tree saved from a differently nested
transform still converts.

Layout changes: conv kernels HWIO -> OIHW, depthwise conv1d [k, 1, C] -> [C, 1, k], Linear
[in, out] -> [out, in], norm scale/offset -> weight/bias.  JAX-only heads that the torch model
does not have (regression_hid, conv_stats_*, ...) are ignored.

PARITY UNPINNED: neither JAX nor a published checkpoint is reachable from this container, so the
name table is derived, not observed.  Every conversion is validated against the torch schema
(all keys present, every shape right); a wrong name fails loudly rather than mis-loading.
"""
from collections import OrderedDict
from typing import Callable, List, Mapping, Tuple

import numpy as np
import torch

from tapnet_b200 import schema


def _canon(name: str) -> str:
  return '/'.join(p for p in name.split('/') if p != '~')


def _suffix(base: str, i: int) -> str:
  """Haiku's auto-numbering of repeated default names: 'x', 'x_1', 'x_2', ..."""
  return base if i == 0 else f'{base}_{i}'


def _conv(w):        # HWIO -> OIHW
  return np.transpose(w, (3, 2, 0, 1))


def _dwconv1d(w):    # [k, 1, C] -> [C, 1, k]
  return np.transpose(w, (2, 1, 0))


def _linear(w):      # [in, out] -> [out, in]
  return np.transpose(w, (1, 0))


def _vec(w):
  return np.reshape(w, (-1,))


_INV = {_conv: lambda w: np.transpose(w, (2, 3, 1, 0)), _dwconv1d: _dwconv1d, _linear: _linear,
        _vec: _vec}


def key_table(pyramid_level: int = 1, extra_convs: bool = True, num_mixer_blocks: int = 12
              ) -> List[Tuple[str, str, str, Callable]]:
  """[(torch key, canonical Haiku module, Haiku parameter, layout transform)] in schema order."""
  del pyramid_level  # the layout does not depend on it (only linear's input width does)
  t = []
  t.append(('resnet_torch.initial_conv.weight', 'tapir/resnet/initial_conv', 'w', _conv))
  for g, nb in enumerate(schema.BLOCKS_PER_GROUP):
    for b in range(nb):
      tp = f'resnet_torch.block_groups.{g}.blocks.{b}.'
      hp = f'tapir/resnet/block_group_{g}/block_{b}/'
      if b == 0:
        t.append((tp + 'proj_conv.weight', hp + 'shortcut_conv', 'w', _conv))
      t.append((tp + 'bn_0.weight', hp + 'instancenorm_0', 'scale', _vec))
      t.append((tp + 'bn_0.bias', hp + 'instancenorm_0', 'offset', _vec))
      t.append((tp + 'conv_0.weight', hp + 'conv_0', 'w', _conv))
      t.append((tp + 'conv_1.weight', hp + 'conv_1', 'w', _conv))
      t.append((tp + 'bn_1.weight', hp + 'instancenorm_1', 'scale', _vec))
      t.append((tp + 'bn_1.bias', hp + 'instancenorm_1', 'offset', _vec))
  head = (('hid1', 'cost_volume_regression_1', _conv), ('hid2', 'cost_volume_regression_2', _conv),
          ('hid3', 'cost_volume_occlusion_1', _conv), ('hid4', 'cost_volume_occlusion_2', _linear),
          ('occ_out', 'occlusion_out', _linear))
  for tname, hname, tf in head:
    t.append((f'torch_cost_volume_track_mods.{tname}.weight', f'tapir/{hname}', 'w', tf))
    t.append((f'torch_cost_volume_track_mods.{tname}.bias', f'tapir/{hname}', 'b', _vec))
  mx, hm = 'torch_pips_mixer.', 'tapir/pips_mlp_mixer/'
  t.append((mx + 'linear.weight', hm + 'linear', 'w', _linear))
  t.append((mx + 'linear.bias', hm + 'linear', 'b', _vec))
  t.append((mx + 'layer_norm.weight', hm + 'layer_norm', 'scale', _vec))
  t.append((mx + 'linear_1.weight', hm + 'linear_1', 'w', _linear))
  t.append((mx + 'linear_1.bias', hm + 'linear_1', 'b', _vec))
  for i in range(num_mixer_blocks):
    tp, hp = f'{mx}blocks.{i}.', hm + _suffix('block', i) + '/'
    t.append((tp + 'layer_norm.weight', hp + 'layer_norm', 'scale', _vec))
    t.append((tp + 'mlp1_up.weight', hp + 'mlp1_up', 'w', _dwconv1d))
    t.append((tp + 'mlp1_up.bias', hp + 'mlp1_up', 'b', _vec))
    t.append((tp + 'mlp1_up_1.weight', hp + 'mlp1_up_1', 'w', _dwconv1d))
    t.append((tp + 'mlp1_up_1.bias', hp + 'mlp1_up_1', 'b', _vec))
    t.append((tp + 'layer_norm_1.weight', hp + 'layer_norm_1', 'scale', _vec))
    t.append((tp + 'conv_channels_mixer.mlp2_up.weight', hp + 'mlp2_up', 'w', _linear))
    t.append((tp + 'conv_channels_mixer.mlp2_up.bias', hp + 'mlp2_up', 'b', _vec))
    t.append((tp + 'conv_channels_mixer.mlp2_down.weight', hp + 'mlp2_down', 'w', _linear))
    t.append((tp + 'conv_channels_mixer.mlp2_down.bias', hp + 'mlp2_down', 'b', _vec))
  if extra_convs:
    for i in range(schema.NUM_EXTRA_CONV_BLOCKS):
      tp, hp = f'extra_convs.blocks.{i}.', 'tapir/extra_convs/'
      t.append((tp + 'layer_norm.weight', hp + _suffix('layer_norm', i), 'scale', _vec))
      t.append((tp + 'layer_norm.bias', hp + _suffix('layer_norm', i), 'offset', _vec))
      t.append((tp + 'conv.weight', hp + _suffix('conv2_d', 2 * i), 'w', _conv))
      t.append((tp + 'conv.bias', hp + _suffix('conv2_d', 2 * i), 'b', _vec))
      t.append((tp + 'conv_1.weight', hp + _suffix('conv2_d', 2 * i + 1), 'w', _conv))
      t.append((tp + 'conv_1.bias', hp + _suffix('conv2_d', 2 * i + 1), 'b', _vec))
  return t


def infer_model_kwargs(params: Mapping[str, Mapping[str, np.ndarray]]) -> dict:
  """TAPIR constructor arguments implied by a Haiku tree: extra_convs present?  mixer input
  width -> pyramid_level (models/tapir_model.py:378-390)."""
  canon = {_canon(k): v for k, v in params.items()}
  extra = any(k.startswith('tapir/extra_convs/') for k in canon)
  din = int(np.shape(canon['tapir/pips_mlp_mixer/linear']['w'])[0])
  levels = (din - 4 - schema.HIRES_DIM - schema.LOWRES_DIM) // (schema.PATCH * schema.PATCH) - 2
  if schema.mixer_input_dim(levels) != din or levels < 0:
    raise ValueError(f'unexpected mixer input width {din}')
  return dict(pyramid_level=levels, extra_convs=extra)


def convert_haiku_params(params: Mapping[str, Mapping[str, np.ndarray]], pyramid_level=None,
                         extra_convs=None, num_mixer_blocks: int = 12
                         ) -> 'OrderedDict[str, torch.Tensor]':
  """Haiku parameter tree -> torch state dict (fp32), validated against the schema."""
  inferred = infer_model_kwargs(params)
  pyramid_level = inferred['pyramid_level'] if pyramid_level is None else pyramid_level
  extra_convs = inferred['extra_convs'] if extra_convs is None else extra_convs
  canon = {_canon(k): v for k, v in params.items()}
  want = schema.state_dict_schema(pyramid_level, extra_convs, num_mixer_blocks)
  out = OrderedDict()
  missing = []
  for tkey, module, pname, tf in key_table(pyramid_level, extra_convs, num_mixer_blocks):
    if module not in canon or pname not in canon[module]:
      missing.append(f'{module}:{pname}')
      continue
    arr = tf(np.asarray(canon[module][pname], dtype=np.float32))
    if tuple(arr.shape) != tuple(want[tkey]):
      raise ValueError(f'{module}:{pname} converts to shape {tuple(arr.shape)}, torch key '
                       f'{tkey} needs {tuple(want[tkey])}')
    out[tkey] = torch.from_numpy(np.ascontiguousarray(arr))
  if missing:
    raise KeyError('Haiku tree lacks ' + ', '.join(missing[:8]) +
                   (f' ... ({len(missing)} in total)' if len(missing) > 8 else ''))
  assert list(out.keys()) == list(want.keys())
  return out


def load_jax_checkpoint(path: str, **kwargs) -> 'OrderedDict[str, torch.Tensor]':
  """`np.load(path, allow_pickle=True).item()['params']` (tapir_clustering.py:923-924) -> torch
  state dict.  Only load files you trust: the format is a pickle."""
  ckpt = np.load(path, allow_pickle=True).item()
  return convert_haiku_params(ckpt['params'] if 'params' in ckpt else ckpt, **kwargs)


def to_haiku_params(state_dict: Mapping[str, torch.Tensor], pyramid_level: int = 1,
                    extra_convs: bool = True, num_mixer_blocks: int = 12, tilde: bool = True):
  """Inverse mapping (torch state dict -> Haiku tree), e.g. to hand torch-trained weights to
  the JAX model.  `tilde` re-inserts the '~' elements of modules built in __init__."""
  def name(canon):
    if not tilde:
      return canon
    parts = canon.split('/')
    out = [parts[0]]
    for i, p in enumerate(parts[1:], 1):
      # children of tapir, resnet, block_group_*, block_* (resnet) are built in __init__
      in_init = (i == 1) or (parts[1] == 'resnet')
      if in_init:
        out.append('~')
      out.append(p)
    return '/'.join(out)
  tree = {}
  for tkey, module, pname, tf in key_table(pyramid_level, extra_convs, num_mixer_blocks):
    arr = _INV[tf](state_dict[tkey].detach().cpu().numpy())
    tree.setdefault(name(module), {})[pname] = np.ascontiguousarray(arr)
  return tree

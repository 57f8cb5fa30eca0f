"""State-dict schema of the TAPIR / BootsTAPIR torch checkpoints.

The published `.pt` files are plain `state_dict`s of `tapnet/torch/tapir_model.py:TAPIR`
(`tapnet/pytorch_live_demo.py:112-114`).  The key names and shapes below are derived from
the module definitions (`tapir_model.py:115-137`, `nets.py:25-89,107-244,247-427`); the test
`tests/test_schema.py` checks them against the real reference module when it is importable.
"""
from collections import OrderedDict
from typing import Tuple

HIRES_DIM = 128
LOWRES_DIM = 256
MIXER_HIDDEN = 512
MIXER_EXPAND = 4
PATCH = 7
GROUP_CHANNELS = (64, 128, 256, 256)
GROUP_STRIDES = (1, 2, 2, 1)
BLOCKS_PER_GROUP = (2, 2, 2, 2)  # tapir_model.py:111 overrides the ctor arg
NUM_EXTRA_CONV_BLOCKS = 5
EXTRA_CONV_MULT = 4


def mixer_input_dim(pyramid_level: int) -> int:
  # tapir_model.py:128-129
  return 4 + HIRES_DIM + LOWRES_DIM + (pyramid_level + 2) * PATCH * PATCH


def mixer_output_dim() -> int:
  return 4 + HIRES_DIM + LOWRES_DIM


def state_dict_schema(
    pyramid_level: int = 1, extra_convs: bool = True, num_mixer_blocks: int = 12
) -> 'OrderedDict[str, Tuple[int, ...]]':
  """Returns key -> shape in torch registration order."""
  s = OrderedDict()
  s['resnet_torch.initial_conv.weight'] = (64, 3, 7, 7)
  cin = 64
  for g, (cout, nb) in enumerate(zip(GROUP_CHANNELS, BLOCKS_PER_GROUP)):
    for b in range(nb):
      p = f'resnet_torch.block_groups.{g}.blocks.{b}.'
      c_in = cin if b == 0 else cout
      if b == 0:
        s[p + 'proj_conv.weight'] = (cout, c_in, 1, 1)
      s[p + 'bn_0.weight'] = (c_in,)
      s[p + 'bn_0.bias'] = (c_in,)
      s[p + 'conv_0.weight'] = (cout, c_in, 3, 3)
      s[p + 'conv_1.weight'] = (cout, cout, 3, 3)
      s[p + 'bn_1.weight'] = (cout,)
      s[p + 'bn_1.bias'] = (cout,)
    cin = cout
  hd = s
  m = 'torch_cost_volume_track_mods.'
  hd[m + 'hid1.weight'] = (16, 1, 3, 3)
  hd[m + 'hid1.bias'] = (16,)
  hd[m + 'hid2.weight'] = (1, 16, 3, 3)
  hd[m + 'hid2.bias'] = (1,)
  hd[m + 'hid3.weight'] = (32, 16, 3, 3)
  hd[m + 'hid3.bias'] = (32,)
  hd[m + 'hid4.weight'] = (16, 32)
  hd[m + 'hid4.bias'] = (16,)
  hd[m + 'occ_out.weight'] = (2, 16)
  hd[m + 'occ_out.bias'] = (2,)
  x = 'torch_pips_mixer.'
  din, dout = mixer_input_dim(pyramid_level), mixer_output_dim()
  h, e = MIXER_HIDDEN, MIXER_HIDDEN * MIXER_EXPAND
  s[x + 'linear.weight'] = (h, din)
  s[x + 'linear.bias'] = (h,)
  s[x + 'layer_norm.weight'] = (h,)
  s[x + 'linear_1.weight'] = (dout, h)
  s[x + 'linear_1.bias'] = (dout,)
  for i in range(num_mixer_blocks):
    p = f'{x}blocks.{i}.'
    s[p + 'layer_norm.weight'] = (h,)
    s[p + 'mlp1_up.weight'] = (e, 1, 3)
    s[p + 'mlp1_up.bias'] = (e,)
    s[p + 'mlp1_up_1.weight'] = (e, 1, 3)
    s[p + 'mlp1_up_1.bias'] = (e,)
    s[p + 'layer_norm_1.weight'] = (h,)
    s[p + 'conv_channels_mixer.mlp2_up.weight'] = (e, h)
    s[p + 'conv_channels_mixer.mlp2_up.bias'] = (e,)
    s[p + 'conv_channels_mixer.mlp2_down.weight'] = (h, e)
    s[p + 'conv_channels_mixer.mlp2_down.bias'] = (h,)
  if extra_convs:
    c = LOWRES_DIM
    for i in range(NUM_EXTRA_CONV_BLOCKS):
      p = f'extra_convs.blocks.{i}.'
      s[p + 'layer_norm.weight'] = (c,)
      s[p + 'layer_norm.bias'] = (c,)
      s[p + 'conv.weight'] = (c * EXTRA_CONV_MULT, c, 3, 3)
      s[p + 'conv.bias'] = (c * EXTRA_CONV_MULT,)
      s[p + 'conv_1.weight'] = (c, c * EXTRA_CONV_MULT, 3, 3)
      s[p + 'conv_1.bias'] = (c,)
  return s

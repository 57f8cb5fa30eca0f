"""CPU: Haiku `.npy` parameter tree -> torch state dict (tapnet_b200/convert.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import synth
from tapnet_b200 import convert, schema


@pytest.mark.parametrize('pyramid_level,extra', [(1, True), (0, False)])
@pytest.mark.parametrize('tilde', [True, False])
def test_round_trip(pyramid_level, extra, tilde):
  sd = synth.make_state_dict(0, pyramid_level, extra)
  tree = convert.to_haiku_params(sd, pyramid_level, extra, tilde=tilde)
  assert convert.infer_model_kwargs(tree) == dict(pyramid_level=pyramid_level, extra_convs=extra)
  back = convert.convert_haiku_params(tree)
  assert list(back.keys()) == list(schema.state_dict_schema(pyramid_level, extra).keys())
  for k, v in sd.items():
    assert torch.equal(back[k], v.float()), k


def test_haiku_names():
  tree = convert.to_haiku_params(synth.make_state_dict(0))
  # prefix and block numbering as in the reference's causal-state keys
  # (tapir_clustering.py:824-847: 'tapir/~/pips_mlp_mixer/block_causal_1', 'block_1_...', 'block_11_...')
  for name in ('tapir/~/pips_mlp_mixer/block/mlp1_up', 'tapir/~/pips_mlp_mixer/block_11/mlp1_up_1',
               'tapir/~/pips_mlp_mixer/linear_1', 'tapir/~/resnet/~/initial_conv',
               'tapir/~/resnet/~/block_group_3/~/block_1/~/instancenorm_1',
               'tapir/~/resnet/~/block_group_1/~/block_0/~/shortcut_conv',
               'tapir/~/extra_convs/conv2_d', 'tapir/~/extra_convs/conv2_d_9',
               'tapir/~/extra_convs/layer_norm_4', 'tapir/~/cost_volume_occlusion_2',
               'tapir/~/occlusion_out'):
    assert name in tree, name
  assert 'tapir/~/pips_mlp_mixer/block_12/mlp1_up' not in tree
  assert tree['tapir/~/resnet/~/initial_conv']['w'].shape == (7, 7, 3, 64)             # HWIO
  assert tree['tapir/~/pips_mlp_mixer/block/mlp1_up']['w'].shape == (3, 1, 2048)       # [k,1,C]
  assert tree['tapir/~/pips_mlp_mixer/linear']['w'].shape == (535, 512)                 # [in,out]
  assert set(tree['tapir/~/pips_mlp_mixer/layer_norm']) == {'scale'}
  assert set(tree['tapir/~/extra_convs/layer_norm']) == {'scale', 'offset'}


def test_layouts_mean_the_same_operation():
  """A Haiku-layout kernel applied with Haiku/lax semantics == the converted kernel in torch."""
  rng = np.random.default_rng(0)
  x = rng.normal(size=(2, 6, 7, 5)).astype(np.float32)          # NHWC
  w = rng.normal(size=(3, 3, 5, 4)).astype(np.float32)          # HWIO
  xp = np.pad(x, ((0, 0), (1, 1), (1, 1), (0, 0)))
  want = np.zeros((2, 6, 7, 4), np.float32)
  for ky in range(3):
    for kx in range(3):                                          # cross-correlation, like lax
      want += np.einsum('nhwi,io->nhwo', xp[:, ky:ky + 6, kx:kx + 7], w[ky, kx])
  got = F.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(convert._conv(w).copy()),
                 padding=1).permute(0, 2, 3, 1).numpy()
  np.testing.assert_allclose(got, want, atol=1e-5)
  # depthwise conv1d, channel multiplier 4: output channel o reads input channel o // 4
  xt = rng.normal(size=(2, 9, 3)).astype(np.float32)            # [N, W, C]
  wd = rng.normal(size=(3, 1, 12)).astype(np.float32)           # [k, 1, C * mult]
  xtp = np.pad(xt, ((0, 0), (1, 1), (0, 0)))
  want = np.zeros((2, 9, 12), np.float32)
  for o in range(12):
    for k in range(3):
      want[:, :, o] += xtp[:, k:k + 9, o // 4] * wd[k, 0, o]
  got = F.conv1d(torch.from_numpy(xt).permute(0, 2, 1), torch.from_numpy(convert._dwconv1d(wd).copy()),
                 padding=1, groups=3).permute(0, 2, 1).numpy()
  np.testing.assert_allclose(got, want, atol=1e-5)
  wl = rng.normal(size=(5, 4)).astype(np.float32)               # [in, out]
  np.testing.assert_allclose(F.linear(torch.from_numpy(x), torch.from_numpy(convert._linear(wl).copy())).numpy(),
                             x @ wl, atol=1e-5)


def test_errors_are_loud():
  tree = convert.to_haiku_params(synth.make_state_dict(0))
  broken = dict(tree)
  del broken['tapir/~/pips_mlp_mixer/block_7/mlp2_up']
  with pytest.raises(KeyError, match='block_7/mlp2_up'):
    convert.convert_haiku_params(broken)
  broken = {k: dict(v) for k, v in tree.items()}
  broken['tapir/~/cost_volume_regression_1']['w'] = np.zeros((3, 3, 2, 16), np.float32)
  with pytest.raises(ValueError, match='hid1'):
    convert.convert_haiku_params(broken)
  extra = dict(tree)
  extra['tapir/~/regression_hid'] = {'w': np.zeros((4, 128), np.float32)}   # JAX-only head: ignored
  assert len(convert.convert_haiku_params(extra)) == 218


def test_load_jax_checkpoint_file_and_model_load(tmp_path):
  from tapnet_b200 import tapir_model
  sd = synth.make_state_dict(0, 0, False)
  path = tmp_path / 'tapir_checkpoint.npy'
  np.save(path, {'params': convert.to_haiku_params(sd, 0, False), 'state': {}}, allow_pickle=True)
  back = convert.load_jax_checkpoint(str(path))
  model = tapir_model.TAPIR(**convert.infer_model_kwargs(convert.to_haiku_params(sd, 0, False)))
  missing, unexpected = model.load_state_dict(back, strict=True)
  assert not missing and not unexpected

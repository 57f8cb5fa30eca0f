import pytest

from oracle import reference_loader
from tapnet_b200 import schema


def test_param_counts():
  s = schema.state_dict_schema()
  n = sum(int(__import__('numpy').prod(v)) for v in s.values())
  assert len(s) == 218 and n == 54699335
  s0 = schema.state_dict_schema(0, False)
  assert len(s0) == 188


@pytest.mark.skipif(not reference_loader.available(), reason='reference not mounted')
@pytest.mark.parametrize('kw', [dict(pyramid_level=1), dict(pyramid_level=0, extra_convs=False)])
def test_schema_matches_reference_module(kw):
  ref = reference_loader.load().TAPIR(**kw).state_dict()
  s = schema.state_dict_schema(kw.get('pyramid_level', 1), kw.get('extra_convs', True))
  assert list(ref.keys()) == list(s.keys())
  for k, v in ref.items():
    assert tuple(v.shape) == s[k], k

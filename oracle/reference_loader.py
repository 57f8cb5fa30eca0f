"""TEST INFRASTRUCTURE - imports the UNMODIFIED reference torch path from /root/reference.

Only usable in the build container (the GPU box has no /root/reference).  Used to pin the
restatement in `oracle/tapir_oracle.py` and to generate `tests/golden/*.npz`
(`oracle/make_golden.py`).  Nothing from the reference is copied into this repo.
"""
import os
import sys

REFERENCE_ROOT = os.environ.get('TAPNET_REFERENCE_ROOT', '/root/reference')
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shims')


def available() -> bool:
  return os.path.isfile(os.path.join(REFERENCE_ROOT, 'tapnet', 'torch', 'tapir_model.py'))


def load():
  """Returns the reference `tapnet.torch.tapir_model` module."""
  if not available():
    raise RuntimeError(f'reference not found under {REFERENCE_ROOT}')
  for p in (REFERENCE_ROOT, _SHIMS):
    if p not in sys.path:
      sys.path.insert(0, p)
  from tapnet.torch import tapir_model  # pylint: disable=g-import-not-at-top
  return tapir_model


def build(state_dict, **kwargs):
  """Reference TAPIR(**kwargs) in eval mode with `state_dict` loaded."""
  m = load().TAPIR(**kwargs)
  m.load_state_dict(state_dict)
  return m.eval()

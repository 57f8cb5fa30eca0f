"""TEST INFRASTRUCTURE - the seeded input generators live in tapnet_b200/synth.py (they are shared
with bench.py); re-exported here for the tests and the golden-file generators."""
from tapnet_b200.synth import *  # noqa: F401,F403
from tapnet_b200.synth import make_queries, make_state_dict, make_video  # noqa: F401

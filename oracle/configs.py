"""Model-section config dictionaries (re-exported from the shared synthetic-workload module so
the oracle, the tests and bench.py agree on the exact same dicts)."""
from egonet_b200.synth import clone, demo_cfgs, make_cfgs, ped_cfgs, tiny_cfgs  # noqa: F401

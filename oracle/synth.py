"""Moved to dahitra_b200/synth.py (the generator is not oracle code); kept as an alias for the pin script and tests."""
from dahitra_b200.synth import *  # noqa: F401,F403
from dahitra_b200.synth import _key_seed  # noqa: F401

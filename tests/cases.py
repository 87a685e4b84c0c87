"""Deterministic synthetic cases shared by the golden generator, the oracle tests and the GPU parity tests.
The builders live in ``workloads.py`` at the repo root (bench.py uses the same ones)."""
from workloads import *  # noqa: F401,F403
from workloads import _st  # noqa: F401

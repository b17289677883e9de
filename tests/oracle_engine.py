"""test double of the CUDA engine: the CPU oracle behind the same interface."""

from oracle.engine import OracleEngine  # noqa: F401

"""Moved to oracle/layer.py (also used by bench.py's CPU arm); kept so the tests' imports stay short."""
from oracle.layer import OracleAdamW, OracleLayer  # noqa: F401

"""TEST-ONLY engine double (lives in oracle/engine_double.py so that bench.py's CPU arm can use it too)."""
from oracle.engine_double import OracleEngine  # noqa: F401

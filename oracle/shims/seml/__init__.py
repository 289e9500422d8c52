"""Name-only stand-in for `seml` (src/scripts/bilevel.py:8). TEST INFRASTRUCTURE ONLY."""
from . import database_utils, misc  # noqa: F401

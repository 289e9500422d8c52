"""Name-only stand-in for torch-geometric 1.3.2. TEST INFRASTRUCTURE ONLY. Nothing here is
called on the hot path."""

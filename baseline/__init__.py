"""Reference-arm material for bench.py (measurement infrastructure, not product code)."""

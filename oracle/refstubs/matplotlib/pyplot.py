"""Test-only stand-in (reference import site: pyvibdmc/analysis/plotter.py:1,14)."""
rcParams = {}

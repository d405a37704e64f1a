"""Test-only empty stand-in so `import matplotlib.pyplot` in the reference succeeds."""

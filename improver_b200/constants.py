"""Constants shared with the reference (improver/constants.py:16,
improver/metadata/constants/__init__.py:9)."""
import numpy as np

DEFAULT_PERCENTILES = (0, 5, 10, 20, 25, 30, 40, 50, 60, 70, 75, 80, 90, 95, 100)
FLOAT_DTYPE = np.float32

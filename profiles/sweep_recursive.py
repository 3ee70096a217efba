"""Sweep the launch knobs of the recursive filter kernels (one process per setting)."""
import os
import subprocess
import sys

here = os.path.dirname(os.path.abspath(__file__))
for env in [{}, {"NBH_RF_YBATCH": "16"}, {"NBH_RF_YBATCH": "4"}, {"NBH_RF_YTHREADS": "64"},
            {"NBH_RF_YTHREADS": "256", "NBH_RF_YBATCH": "16"}]:
    e = dict(os.environ, **env)
    r = subprocess.run([sys.executable, os.path.join(here, "time_recursive.py")], env=e, capture_output=True, text=True)
    print(env, [l for l in r.stdout.splitlines() if l.startswith(("plain", "mask"))] or r.stderr[-300:], flush=True)

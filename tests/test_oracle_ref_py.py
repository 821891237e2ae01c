"""oracle/ref_py.py: the unmodified Python reference as a timed CPU baseline.  Loaded from the mounted source tree
in the build container and from the byte-code compiled out of it (oracle/_ref/refpy/, what travels to the GPU box);
both must behave as the reference's own tests say (tests/test_envs.py:108-137, tests/test_standardization.py)."""
import subprocess
import sys
from pathlib import Path

import pytest

from oracle import ref_py

REPO = Path(__file__).resolve().parent.parent

CHECK = """
import sys
sys.path.insert(0, {repo!r})
from pathlib import Path
import numpy as np
import oracle.ref_py as R
if {force_bytecode!r}:
    R.REF = Path("/nonexistent")
SimpleGrid, Est, BensLog, Norm, Welford, origin = R.load_reference()
assert origin == ("bytecode" if {force_bytecode!r} else "source"), origin
table = {{1: ((5, 6), -1), 2: ((6, 5), -1), 3: ((4, 5), 0), 4: ((5, 4), 0)}}
for a, (xy, rew) in table.items():
    env = SimpleGrid(np.array([5, 5], np.int32), np.array([0, 0], np.int32), size=10)
    env.reset()
    obs, r, done, trunc, info = env.step(a)
    assert tuple(int(v) for v in obs) == xy and r == rew and not done and not trunc
w = Welford()
w.update(1000); w.update(2000)
assert w.mean == 1500 and abs(w.std - 707.10678) < 1e-4
rate, n = R.time_simplegrid(0.05)
assert rate > 1000 and n >= 1
rate, n = R.time_pipeline(0.05)
assert rate > 1000
print("ok", origin)
"""


@pytest.mark.parametrize("force_bytecode", [False, True])
def test_reference_loads_and_behaves(force_bytecode):
    if not ref_py.available():
        pytest.skip("reference neither mounted nor byte-compiled")
    if force_bytecode and not all(ref_py._bc_path(rel).exists() for rel in ref_py.FILES):
        pytest.skip("byte-code not built (make -C oracle ref_py)")
    if not force_bytecode and not ref_py.REF.exists():
        pytest.skip("reference not mounted")
    out = subprocess.run([sys.executable, "-c", CHECK.format(repo=str(REPO), force_bytecode=force_bytecode)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().endswith("bytecode" if force_bytecode else "source")

"""Config 5 consumer: the PPO rollout example runs end to end on a tiny problem and reports the env
fraction of the step time."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


def test_ppo_rollout_example_runs():
    out = subprocess.run([sys.executable, str(REPO / "examples" / "ppo_rollout.py"), "--envs", "96", "--horizon", "130",
                          "--iters", "2", "--minibatches", "2", "--dtype", "fp32"], capture_output=True, text=True,
                         timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    res = json.loads(out.stdout.strip().splitlines()[-1])
    assert res["env_error_flags"] == 0 and len(res["iters"]) == 2
    assert 0.0 < res["iters"][-1]["env_fraction_of_rollout_step"] < 1.0
    assert res["iters"][-1]["loss"] == res["iters"][-1]["loss"]  # not NaN

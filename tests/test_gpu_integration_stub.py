"""INTEGRATION.md's ctypes stub is executable documentation: run it (at a small size) on the GPU."""
import re
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


def test_integration_md_stub_runs():
    import torch

    from rad_team_b200 import _lib

    md = (REPO / "INTEGRATION.md").read_text()
    block = re.search(r"## 2\. The stub.*?```python\n(.*?)```", md, flags=re.S).group(1)
    block = block.replace('C.CDLL("librt_b200.so")', f'C.CDLL("{_lib.lib_path()}")').replace("n_envs=65536", "n_envs=512")
    ns = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    torch.cuda.synchronize()
    E, A, G = ns["E"], ns["A"], ns["G"]
    assert (E, A, G) == (512, 3, 32)
    maps, xy, reward = ns["obs_maps"], ns["obs_xy"], ns["reward"]
    assert int(maps[:, 0].sum()) >= E and float(maps.max()) <= 1.0            # agents marked, values normalised
    assert int(xy.min()) >= 0 and int(xy.max()) < G and set(reward.unique().tolist()) <= {-1, 0, 1}
    assert ns["L"].rt_env_destroy(ns["env"]) == 0

"""The drop-in boundary: librt_b200.so loads without a GPU, exports every symbol include/rt_b200.h
declares, and fails loudly (never falls back to a CPU path) when there is no device.  CPU only."""
import ctypes as C
import re
from pathlib import Path

import pytest

from rad_team_b200 import _lib

HEADER = Path(__file__).resolve().parent.parent / "include" / "rt_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    L = C.CDLL(str(_lib.lib_path()))
    for name in declared_symbols():
        assert hasattr(L, name), name
    assert _lib.lib().rt_abi_version() == 1


def test_cfg_struct_layout_matches_header():
    # 10 x int32, int64, uint64, 7 x double, naturally aligned
    assert C.sizeof(_lib.EnvCfg) == 10 * 4 + 8 + 8 + 7 * 8
    assert _lib.EnvCfg.env_id_base.offset == 40 and _lib.EnvCfg.cell_pitch_cm.offset == 56


def test_bad_config_rejected_before_touching_cuda():
    L = _lib.lib()
    h = C.c_void_p()
    for bad in (dict(n_envs=0), dict(n_agents=9), dict(grid=1), dict(max_steps=0), dict(max_steps=1, n_agents=1),
                dict(max_steps=40000, n_agents=2), dict(poly_max=6), dict(poly_min=2, poly_max=1),
                dict(cell_pitch_cm=0.0), dict(lognorm_step=0), dict(grid=256), dict(src_lo=-1.0),
                dict(src_lo=5.0, src_hi=1.0), dict(bkg_lo=2.0, bkg_hi=1.0)):
        kw = dict(n_envs=4, n_agents=1, grid=8, max_steps=10, lognorm_step=2, cell_pitch_cm=100.0, min_dist_cm=50.0)
        kw.update(bad)
        c = _lib.EnvCfg(**kw)
        assert L.rt_env_create(C.byref(c), C.byref(h)) == _lib.RT_ERR_BAD_ARG, bad
    assert L.rt_env_create(None, C.byref(h)) == _lib.RT_ERR_BAD_ARG
    assert L.rt_env_step(None, None, None, None, None, None, None) == _lib.RT_ERR_BAD_ARG


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _lib.lib()
    h = C.c_void_p()
    c = _lib.EnvCfg(n_envs=4, n_agents=1, grid=8, max_steps=10, lognorm_step=2, cell_pitch_cm=100.0, min_dist_cm=50.0)
    assert L.rt_env_create(C.byref(c), C.byref(h)) == _lib.RT_ERR_CUDA
    assert L.rt_welford_create(4, C.byref(h)) == _lib.RT_ERR_CUDA
    assert L.rt_moments_create(C.byref(h)) == _lib.RT_ERR_CUDA
    assert L.rt_estimator_create(1, 4, 4, C.byref(h)) == _lib.RT_ERR_CUDA
    from rad_team_b200.envs import BatchedRadSearchEnv, SimpleGrid
    from rad_team_b200.math_utils import WelfordsOnlineStandardization

    with pytest.raises(_lib.RtError):
        BatchedRadSearchEnv(n_envs=2)
    with pytest.raises(_lib.RtError):
        SimpleGrid([5, 5], [0, 0])
    with pytest.raises(_lib.RtError):
        WelfordsOnlineStandardization()


def test_product_never_imports_the_oracle():
    """Nothing shipped may import, include, link or dlopen anything under oracle/."""
    pkg = Path(_lib.__file__).resolve().parent
    pat = re.compile(r"(import\s+oracle|from\s+oracle|from\s+\.+oracle|oracle/|rt_oracle|orc_[a-z_]+\s*\()")
    files = [f for ext in ("*.py", "*.cu", "*.cuh", "*.h") for f in pkg.rglob(ext)] + [HEADER]
    assert len(files) > 8
    for f in files:
        assert not pat.search(f.read_text()), f


def test_header_is_plain_c_and_the_integration_stub_compiles(tmp_path):
    """include/rt_b200.h must be consumable by a C99 compiler (the drop-in boundary is a C ABI), and the C snippet of
    INTEGRATION.md section 5 must compile against it."""
    import shutil
    import subprocess

    gcc = shutil.which("gcc") or "/usr/bin/gcc"
    src = tmp_path / "stub.c"
    src.write_text(r'''
#include "rt_b200.h"
int run_rollout_close(int rank, int world, rt_env* env, const float* readings_dev, long long n, double* ep_dev,
                      double* ep_sum_dev, void* stream) {
  rt_comm* comm;
  rt_moments *local, *global;
  int rc = rt_comm_create_shm("/my_job_1234", rank, world, &comm);
  if (rc != RT_OK) return rc;
  rt_moments_create(&local);
  rt_moments_create(&global);
  rt_moments_update(local, readings_dev, n, stream);
  rt_env_episode_stats(env, ep_dev, stream);
  rc = rt_stats_allgather_merge(comm, local, global, ep_dev, ep_sum_dev, 0, stream);
  rt_comm_disconnect(comm);
  rt_comm_destroy(comm);
  return rc;
}
int main(void) {
  rt_env_cfg cfg = {0};
  size_t bytes = 0;
  cfg.n_envs = 4; cfg.n_agents = 1; cfg.grid = 8; cfg.max_steps = 10; cfg.lognorm_step = 2;
  cfg.cell_pitch_cm = 100.0; cfg.min_dist_cm = 50.0;
  (void)bytes;
  return (int)sizeof(cfg) == 96 && RT_ABI_VERSION == 1 && RT_COMM_HANDLE_BYTES == 64 ? 0 : 1;
}
''')
    out = subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", str(HEADER.parent), "-c",
                          str(src), "-o", str(tmp_path / "stub.o")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr

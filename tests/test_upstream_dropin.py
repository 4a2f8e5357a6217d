"""The real upstream tree on the kernels: `install()` rebinds the names upstream's own modules
resolve, signatures equal `inspect.signature` of upstream's functions, guarded patches keep upstream's classes, and -- on
a GPU -- the unmodified models produce the same disparity pyramid patched and unpatched.  Needs an upstream checkout
($DSSMD_REFERENCE, /root/reference in the build container, or <repo>/_upstream staged by scripts/stage_upstream.sh);
skipped where there is none (the GPU box of the round-end run).  Runs in a subprocess: importing upstream claims the
top-level names `models`, `utils`, `losses`."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "upstream_dropin.py")


def _run(*flags, timeout=1200):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, SCRIPT, *flags], capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    try:
        doc = json.loads(r.stdout)
    except ValueError:
        raise AssertionError(f"upstream_dropin.py {flags}: rc {r.returncode}\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}")
    if "skipped" in doc:
        pytest.skip(doc["skipped"])
    return r.returncode, doc


def test_install_on_real_upstream_modules():
    rc, doc = _run("--binding")
    failed = [c for c in doc["binding"]["checks"] if not c["ok"]]
    assert not failed, failed
    assert rc == 0
    names = {c["name"] for c in doc["binding"]["checks"]}
    for m in ("models.SSMDNet", "models.DSSMD", "models.DSSMD_new", "models.SDNet"):
        assert f"rebound:{m}.build_corr" in names
    for k in ("build_corr", "disp_warp", "LRwarp_error", "CostVolume.forward"):
        assert f"signature:{k}" in names


@pytest.mark.gpu
def test_real_upstream_models_give_the_same_pyramid():
    rc, doc = _run("--gpu", "--iters", "2", timeout=3000)
    bad = [m for m in doc["gpu"]["models"] if not m["ok"]]
    assert not bad, bad
    assert doc["gpu"]["train_step"]["ok"], doc["gpu"]["train_step"]
    assert rc == 0

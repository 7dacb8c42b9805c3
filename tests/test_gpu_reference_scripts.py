"""The reference's OWN demo.py, unmodified, executed end to end against forge_b200 through compat/ (BASELINE.json north_star:
"kubric_train_*.py and demo.py run unchanged").  Needs the staged reference tarball (tools/stage_reference.sh ->
baseline/_ref/forge_reference.tar.gz, git-ignored) or a checkout at $FORGE_REFERENCE; skipped otherwise.  The script predicts
poses, runs its 3 x 2001 refinement iterations through rotate -> fuse -> heads -> render -> backward, renders 28 novel views
per case and writes GIFs; weights are random-init synthetic checkpoints with the reference's state_dict keys."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAVE = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "forge_reference.tar.gz")) or \
    os.path.isdir(os.path.join(os.environ.get("FORGE_REFERENCE", "/root/reference"), "models"))


@pytest.mark.skipif(not HAVE, reason="no reference checkout / staged tarball")
def test_reference_demo_py_runs_unchanged():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_reference_script.py"), "demo"],
                         capture_output=True, text=True, timeout=1500)
    last = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    summary = json.loads(last[-1])
    assert summary["rc"] == 0
    assert summary["gifs"] == ["0_0.gif", "1_0.gif", "2_0.gif"]          # one 360-degree render per demo case

"""The parked kernel variants (experiments/*.patch) must keep applying to the tree."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCHES = sorted(f for f in os.listdir(os.path.join(ROOT, "experiments")) if f.endswith(".patch"))


@pytest.mark.parametrize("patch", PATCHES)
def test_patch_applies(patch):
    res = subprocess.run(["git", "apply", "--check", os.path.join("experiments", patch)], cwd=ROOT,
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-1500:]

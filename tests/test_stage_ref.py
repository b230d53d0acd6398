"""oracle/stage_ref.py: the verbatim staged copy of the reference (test / measurement infrastructure for the GPU box)."""
import os

import pytest

from oracle import ref_harness as RH
from oracle import stage_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_staged_copy_is_verbatim_ignored_by_git_and_found_by_the_harness():
    if not os.path.isdir(os.path.join(stage_ref.SRC, "models")) and stage_ref.staged_root() is None:
        pytest.skip("reference tree neither mounted nor staged")
    dst = stage_ref.stage()
    assert dst is not None and os.path.isfile(os.path.join(dst, "models", "ppst_model.py"))
    if os.path.isdir(os.path.join(stage_ref.SRC, "models")):
        assert stage_ref._same_tree(stage_ref.SRC, dst)            # byte for byte
    assert RH.available() and os.path.isdir(os.path.join(RH.REF_ROOT, "models"))
    with open(os.path.join(ROOT, ".gitignore")) as f:
        assert "oracle/_ref/" in f.read()                          # the reference's sources never enter this repository's history
    ign = os.path.join(ROOT, ".gpurunignore")
    if os.path.exists(ign):
        with open(ign) as f:
            assert "oracle/_ref" not in f.read()                   # ... but they travel to the GPU box with the snapshot

"""Stage the UNMODIFIED reference tree for use on the GPU box.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  /root/reference exists in the build container but not on the GPU box; this
script copies it verbatim (Python sources and the two CUDA op sources, ~0.8 MB) to oracle/_ref/reference/, which is
git-ignored (the reference's sources never enter this repository's history) but travels with the gpurun snapshot, like
the built .so files.  Nothing under ppst_b200/ imports it.  Users:
  * bench.py --impl reference / the cpu_baseline leg: the reference's own modules on the host cores
    (use_custom_kernel = False -> upfirdn2d_native / F.leaky_relu), `kind: "reference"`;
  * tests/test_gpu_reference_modules.py: reference ConvLayer / ResBlock / Discriminator / SimpleSwappingEvaluator running
    on top of ppst_b200.install();
  * tools/ref_gpu_ops.py: the reference's own JIT CUDA ops (upfirdn2d_kernel.cu, fused_bias_act_kernel.cu) timed next
    to ours on the B200.

    python oracle/stage_ref.py            # copy /root/reference -> oracle/_ref/reference (no-op if the source is absent)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref", "reference")


def staged_root():
    """The reference tree to import from: the mounted original if present, else the staged copy, else None."""
    if os.path.isdir(os.path.join(SRC, "models")):
        return SRC
    if os.path.isdir(os.path.join(DST, "models")):
        return DST
    return None


def stage(force=False):
    if not os.path.isdir(os.path.join(SRC, "models")):
        return DST if os.path.isdir(os.path.join(DST, "models")) else None
    if os.path.isdir(DST):
        if not force and _same_tree(SRC, DST):
            return DST
        shutil.rmtree(DST)
    os.makedirs(os.path.dirname(DST), exist_ok=True)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    with open(os.path.join(os.path.dirname(DST), "README"), "w") as f:
        f.write("Verbatim copy of %s made by oracle/stage_ref.py (git-ignored; test / measurement infrastructure only).\n" % SRC)
    return DST


def _same_tree(a, b):
    for dirpath, dirs, files in os.walk(a):
        dirs[:] = [d for d in dirs if d not in (".git", "__pycache__")]
        for f in files:
            if f.endswith(".pyc"):
                continue
            pa = os.path.join(dirpath, f)
            pb = os.path.join(b, os.path.relpath(pa, a))
            if not os.path.exists(pb) or os.path.getsize(pa) != os.path.getsize(pb):
                return False
            with open(pa, "rb") as fa, open(pb, "rb") as fb:
                if fa.read() != fb.read():
                    return False
    return True


if __name__ == "__main__":
    out = stage(force="--force" in sys.argv)
    print(out or "reference tree not available: nothing staged")

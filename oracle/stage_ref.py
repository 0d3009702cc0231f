"""Stage the UNMODIFIED reference hot-path files as oracle/_ref/ so the benchmark's reference arm can run the reference
itself (not the port) on the GPU box's host cores.

    python -m oracle.stage_ref          # in the build container, where /root/reference exists (also run by build())

TEST / BENCH INFRASTRUCTURE.  oracle/_ref/ is a build artefact: git-ignored (the reference's sources never enter this
repository's history), not gpurun-ignored (it travels to the GPU box with the snapshot like the built .so files).  Files are
copied byte for byte from where they lie under /root/reference; MANIFEST.json records the sha256 of each.  Only the files
on the hot path are staged: models.py, losses.py, postprocess.py, modules/*.  The package never
imports oracle/_ref; only bench.py's cpu_baseline / `--impl reference` legs and the golden generator do.
"""
import hashlib
import json
import os
import shutil

REF_SRC = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["models.py", "losses.py", "postprocess.py", "modules/__init__.py", "modules/basic.py", "modules/resnet.py",
         "modules/segmentation_body.py", "modules/segmentation_head.py"]


def stage(verbose=False):
    if not os.path.isdir(REF_SRC):
        return None
    man = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(DST, "src", rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            man[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_SRC, "sha256": man}, f, indent=1)
    if verbose:
        print("staged", len(man), "reference files into", DST)
    return DST


def staged_src():
    """Path of the staged reference sources, or None."""
    p = os.path.join(DST, "src")
    return p if os.path.isfile(os.path.join(p, "models.py")) else None


if __name__ == "__main__":
    if stage(verbose=True) is None:
        print("reference not present at", REF_SRC)

"""TEST INFRASTRUCTURE ONLY -- recipe that makes the unmodified reference travel to the GPU box.

    python oracle/make_ref.py            (also run by __graft_entry__.build() when /root/reference exists)

The reference (QianWangX/VidSeg_diffusion @ 3e96366, pure Python) cannot be pip-installed here: its build backend
(hatchling) and most of its declared dependencies are absent and there is no index.  ``/root/reference`` does not
exist on the GPU box, so the two packages the hot path lives in -- ``sgm/`` and ``scripts/`` -- are mirrored, byte for
byte, into ``oracle/_ref/`` (git-ignored: reference sources never enter this repository's history; NOT
gpurun-ignored: the directory ships with the snapshot like a built .so).  ``oracle/ref_import.py`` then imports the
reference's own modules from there, with inert stubs for the third-party imports this image lacks; ``bench.py --impl
reference`` and the ``cpu_baseline`` leg time them, the full-size parity tests compare against them.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("VIDSEG_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
PACKAGES = ("sgm", "scripts")


def make_ref(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "sgm")):
        if verbose:
            print(f"make_ref: {SRC} not present (GPU box): keeping {DST} as shipped", file=sys.stderr)
        return os.path.isdir(os.path.join(DST, "sgm"))
    ignore = shutil.ignore_patterns("__pycache__", "*.pyc", "*.mp4", "*.png", "*.jpg", "*.gif")
    for pkg in PACKAGES:
        dst = os.path.join(DST, pkg)
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(SRC, pkg), dst, ignore=ignore)
    n = sum(len(fs) for _, _, fs in os.walk(DST))
    if verbose:
        print(f"make_ref: mirrored {', '.join(PACKAGES)} from {SRC} into {DST} ({n} files)", file=sys.stderr)
    return True


if __name__ == "__main__":
    sys.exit(0 if make_ref() else 1)

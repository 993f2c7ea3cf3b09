"""Recipe for `oracle/_ref/`: the UNMODIFIED reference package, as a built artefact next to the oracle.

TEST / MEASUREMENT INFRASTRUCTURE ONLY.  The reference (latentnetworks/vimure) is pure Python, so "building" it is an
install: its package directory is copied as-is from `/root/reference/src/python/vimure` into `oracle/_ref/vimure`
(git-ignored: no reference source enters the history; not gpurun-ignored: it travels to the GPU box like a built `.so`).
It is imported there only by `bench.py`'s CPU legs (`cpu_baseline` / `--impl reference`: configs 1 and 2 timed on the
box's host cores) through the `oracle/shims` stand-ins for its two missing third-party modules (sktensor, igraph).
Nothing under `vimure_b200/` imports it.

    python oracle/build_ref.py      # in the build container (the only place /root/reference exists)
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/src/python/vimure"
DST = os.path.join(HERE, "_ref", "vimure")


def build_ref(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/_ref: /root/reference is not present here; keeping whatever is already installed")
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, DST, ignore=shutil.ignore_patterns("__pycache__", "test", "*.pyc"))
    if verbose:
        print("oracle/_ref: installed the reference package from", SRC)
    return True


if __name__ == "__main__":
    sys.exit(0 if build_ref() else 1)

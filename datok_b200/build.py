"""Builds the in-tree native libraries:  python -m datok_b200.build

  datok_b200/libdatok_b200.so    CUDA kernels + C ABI, nvcc -gencode arch=compute_100a,code=sm_100a
  datok_b200/libdatok_corpus.so  synthetic corpus generator (bench / test tooling)
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def build(verbose=False):
    env = dict(os.environ)
    env.setdefault("PATH", "")
    env["PATH"] = "/usr/local/cuda/bin:" + env["PATH"]
    out = subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), "all"], env=env, stdout=subprocess.PIPE,
                         stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode:
        sys.stderr.write(out.stdout)
    if out.returncode:
        raise RuntimeError("building libdatok_b200.so failed")
    for f in ("libdatok_b200.so", "libdatok_corpus.so"):
        if not os.path.exists(os.path.join(HERE, f)):
            raise RuntimeError(f + " was not produced")


if __name__ == "__main__":
    build(verbose=True)

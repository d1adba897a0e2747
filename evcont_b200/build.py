"""Build ``libevcont_b200.so`` in-tree:  ``python -m evcont_b200.build``."""
import os
import subprocess
import sys

CSRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc")


def build(verbose=False, jobs=None):
    jobs = jobs or os.cpu_count() or 4
    cmd = ["make", "-C", CSRC, f"-j{jobs}"]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        sys.stdout.write(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("building libevcont_b200.so failed (see output above)")
    return os.path.join(os.path.dirname(CSRC), "libevcont_b200.so")


if __name__ == "__main__":
    print(build(verbose=True))

"""In-tree build of the CUDA library (sm_100a only).

    python -m flowstate_b200.build [--force]

Produces flowstate_b200/libflowstate_b200.so with nvcc; the .so is git-ignored
but travels to the GPU box with the working tree.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libflowstate_b200.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def nvcc_path():
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    objs = []
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc_path(), *ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
               "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        for flag in os.environ.get("FS_NVCC_FLAGS", "").split():          # development: extra -D switches
            cmd.insert(1, flag)
        if os.environ.get("FS_TC_TIMERS") in ("1", "2"):   # in-kernel timers for scripts/tc_debug.py (development only;
            cmd.insert(1, "-DFS_TC_TIMERS=" + os.environ["FS_TC_TIMERS"])   # 2: slots 11-13 hold a timeline instead)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed on %s" % src)
    link = [nvcc_path(), *ARCH, "-shared", "-o", LIB, *objs]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

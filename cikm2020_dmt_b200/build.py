"""Build `libdmt_b200.so` in-tree with nvcc for sm_100a (no torch headers, no JIT cache)."""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libdmt_b200.so")
STAMP = os.path.join(HERE, "csrc", ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "--expt-relaxed-constexpr", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "-Xptxas", "-v",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    paths = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    paths.append(os.path.join(HERE, "..", "include", "dmt_b200.h"))
    for p in paths:
        with open(p, "rb") as fh:
            h.update(os.path.basename(p).encode())    # not the absolute path: the tree is copied to other boxes
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def source_digest():
    """Digest of the sources next to this file (what `dmt_build_digest()` of a current library returns)."""
    return _digest()


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def _up_to_date(digest):
    if os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            return fh.read().strip() == digest
    return False


def build(force=False, verbose=False):
    """Build if the sources changed.  Safe under `torchrun`: the ranks of one box serialise on a lock file and
    all but the first find the library up to date."""
    digest = _digest()
    if not force and _up_to_date(digest):
        return LIB
    import fcntl
    os.makedirs(os.path.join(HERE, "csrc", "_obj"), exist_ok=True)
    with open(os.path.join(HERE, "csrc", "_obj", ".build_lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and _up_to_date(digest):
                return LIB
            return _build_locked(digest, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


def _build_locked(digest, verbose):
    nvcc = nvcc_path()
    if nvcc is None:
        if os.path.exists(LIB):
            return LIB   # no compiler: abi.load() compares the library's own digest and refuses a stale one
        raise RuntimeError("nvcc not found and %s is missing" % LIB)
    objdir = os.path.join(HERE, "csrc", "_obj")
    os.makedirs(objdir, exist_ok=True)
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + ["-I", os.path.join(HERE, "..", "include"), "-c", src, "-o", obj]
        if os.path.basename(src) == "abi.cu":       # dmt_build_digest(): checked by abi.load()
            cmd += ['-DDMT_BUILD_DIGEST="%s"' % digest]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append("== %s\n%s" % (os.path.basename(src), out))
        failed |= p.returncode != 0
    with open(os.path.join(objdir, "build.log"), "w") as fh:
        fh.write("\n".join(log))
    if failed or verbose:
        sys.stderr.write("\n".join(log))
    if failed:
        raise RuntimeError("nvcc failed; see %s" % os.path.join(objdir, "build.log"))
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
                                                 "-Xcompiler", "-fPIC"]
    subprocess.run(cmd, check=True)
    os.replace(tmp, LIB)                      # atomic: a concurrent loader never sees a half-written library
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

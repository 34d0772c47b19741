"""Build libmmx.so (sm_100a) in-tree:  python -m motionmixerconv_b200.build [--force] [--verbose]

Every ``csrc/*.cu`` translation unit is compiled by nvcc in parallel
(``-gencode arch=compute_100a,code=sm_100a -lineinfo``) and linked into
``motionmixerconv_b200/libmmx.so``.  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmmx.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


_INC = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(src, seen=None):
    """Transitive quoted includes of one translation unit (so editing a header only rebuilds its users)."""
    seen = set() if seen is None else seen
    src = os.path.normpath(src)
    if src in seen or not os.path.exists(src):
        return seen
    seen.add(src)
    with open(src) as f:
        for inc in _INC.findall(f.read()):
            _deps(os.path.join(os.path.dirname(src), inc), seen)
    return seen


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    newest = max(os.path.getmtime(d) for d in _deps(src))
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj, r.stderr


STAMP = LIB + ".stamp"


def _source_hash():
    """Content hash of everything that goes into libmmx.so (sources, headers, flags): the up-to-date check must not depend on
    file mtimes, which a snapshot copy to another machine may not preserve."""
    import hashlib
    h = hashlib.sha256(" ".join(FLAGS).encode())
    files = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(os.path.dirname(HERE), "include", "mmx.h")]
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _stamp_ok(digest):
    try:
        with open(STAMP) as f:
            return os.path.exists(LIB) and f.read().strip() == digest
    except OSError:
        return False


def build(force=False, verbose=False):
    digest = _source_hash()
    if not force and not verbose and _stamp_ok(digest):
        return LIB                      # built from exactly these sources (possibly on another machine)
    import fcntl
    os.makedirs(OBJ, exist_ok=True)
    with open(os.path.join(OBJ, ".lock"), "w") as lock:      # one builder at a time (torchrun starts N ranks at once)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not verbose and _stamp_ok(digest):
            return LIB
        path = _build_locked(force, verbose)
        with open(STAMP, "w") as f:
            f.write(digest + "\n")
        return path


def _build_locked(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = _sources()
    logs = []
    with cf.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [o for o, _ in results]
    logs = [l for _, l in results if l]
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < newest:
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

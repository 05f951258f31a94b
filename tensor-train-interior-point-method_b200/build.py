"""Build libttipm_b200.so (sm_100a) in-tree with nvcc.  Usage: python build.py [--force]"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libttipm_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
OBJ = os.path.join(HERE, "_obj")
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr",
          "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
LFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-lcublas", "-lcusolver"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(HERE, "..", "include", "ttipm.h")]


def obj_of(src):
    return os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")


def stale_objects(force=False):
    """sources whose object file is missing or older than the source or any header (every source includes the headers)"""
    hdr_t = max(os.path.getmtime(h) for h in headers())
    out = []
    for src in sources():
        o = obj_of(src)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(src), hdr_t):
            out.append(src)
    return out


def stale():
    if not os.path.exists(OUT) or stale_objects():
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(obj_of(s)) > t for s in sources())


def build(force=False, verbose=False):
    """one object per source (compiled in parallel, only the stale ones), then one link"""
    if not force and not stale():
        return OUT
    from concurrent.futures import ThreadPoolExecutor
    os.makedirs(OBJ, exist_ok=True)
    todo = stale_objects(force)

    def compile_one(src):
        cmd = [NVCC] + CFLAGS + ["-c", src, "-o", obj_of(src)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, " ".join(cmd) + "\n" + res.stdout + res.stderr, res.returncode

    logs, failed = [], False
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1, max(1, len(todo)))) as ex:
        for src, log, rc in ex.map(compile_one, todo):
            logs.append(log)
            failed = failed or rc != 0
    if not failed:
        cmd = [NVCC] + LFLAGS + ["-o", OUT] + [obj_of(s) for s in sources()]
        res = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(" ".join(cmd) + "\n" + res.stdout + res.stderr)
        failed = res.returncode != 0
    log = "\n".join(logs)
    with open(os.path.join(HERE, "build.log"), "a" if (todo and len(todo) < len(sources())) else "w") as f:
        f.write(log)
    if failed:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libttipm_b200.so")
    if verbose:
        print(log)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

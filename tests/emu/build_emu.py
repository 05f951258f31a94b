"""TEST INFRASTRUCTURE: compile the kernel sources as plain C++ (-DTTIPM_EMU) into
tests/emu/_build/libttipm_emu.so so the CPU-only test tier can execute the real kernel
code (threads = OS threads, see csrc/emu.h).  Never used by the product path."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
CSRC = os.path.join(ROOT, "tensor-train-interior-point-method_b200", "csrc")
OUTDIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUTDIR, "libttipm_emu.so")


def build(force=False):
    srcs = sorted(glob.glob(os.path.join(CSRC, "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(ROOT, "include", "ttipm.h")]
    if not force and os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(OUTDIR, exist_ok=True)
    objs = []
    procs = []
    for s in srcs:
        o = os.path.join(OUTDIR, os.path.basename(s) + ".o")
        objs.append(o)
        procs.append(subprocess.Popen(["g++", "-x", "c++", "-std=c++17", "-O1", "-g", "-fPIC", "-DTTIPM_EMU", "-pthread",
                                       "-Wno-unknown-pragmas", "-c", s, "-o", o], stderr=subprocess.PIPE, text=True))
    bad = False
    for p in procs:
        err = p.communicate()[1]
        if p.returncode != 0:
            sys.stderr.write(err)
            bad = True
    if bad:
        raise RuntimeError("emu build failed")
    subprocess.check_call(["g++", "-shared", "-pthread", "-o", OUT] + objs)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

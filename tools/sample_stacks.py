"""Poor man's sampling profiler for long end-to-end runs on the GPU box: run a script with a stack dump of the main
thread every few seconds (faulthandler), then tools/sample_stacks.py --summarise <log> counts the innermost frames.

  python tools/sample_stacks.py 5 gpurun_out/stacks.log oracle/ref_harness/run_dropin_ipm.py maxcut 13 2 83 --cuda --skip-ref
"""
import collections
import faulthandler
import re
import runpy
import sys


def summarise(path):
    samples, cur = [], []
    for line in open(path, errors="replace"):
        if line.startswith("Thread") or line.startswith("Stack"):
            if cur:
                samples.append(cur)
            cur = []
        m = re.match(r'\s+File "(.*?)", line (\d+) in (\S+)', line)
        if m:
            cur.append((m.group(1).split("/")[-1], int(m.group(2)), m.group(3)))
    if cur:
        samples.append(cur)
    print(len(samples), "samples")
    for depth, title in ((1, "innermost frame"), (3, "innermost 3 frames")):
        c = collections.Counter(" <- ".join(f"{f}:{fn}:{ln}" for f, ln, fn in s[:depth]) for s in samples if s)
        print("--", title)
        for k, v in c.most_common(25):
            print(f"{v:5d}  {k}")
    owners = collections.Counter()
    for s in samples:
        for f, ln, fn in s:
            if f in ("tt_ipm.py", "tt_als.py", "eigen.py", "tt.py", "amen.py", "als_product.py", "devtt.py"):
                owners[f"{f}:{fn}"] += 1
                break
    print("-- first frame inside the path's modules")
    for k, v in owners.most_common(25):
        print(f"{v:5d}  {k}")


if __name__ == "__main__":
    if sys.argv[1] == "--summarise":
        summarise(sys.argv[2])
    else:
        period, log = float(sys.argv[1]), open(sys.argv[2], "w")
        faulthandler.dump_traceback_later(period, repeat=True, file=log)
        sys.argv = sys.argv[3:]
        runpy.run_path(sys.argv[0], run_name="__main__")

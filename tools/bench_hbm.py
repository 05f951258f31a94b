"""Memory-bound TT primitives at sizes far above the 126 MB L2 (north star: >= 60 % of HBM bandwidth):
algorithmic bytes (8 x elements read + written) / CUDA-event time, against MEASURED_PEAKS.json's hbm_gbs."""
import json
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path[:0] = [os.path.join(ROOT, "tensor-train-interior-point-method_b200")]
from ttipm_b200 import get_runtime, kernels as K  # noqa: E402


def time_ms(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    rt = get_runtime()
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    rnd = lambda *s: torch.randn(*s, dtype=torch.float64, device=dev, generator=g)
    out = []

    def rec(name, nbytes, fn, note=""):
        ms = time_ms(fn)
        gbs = nbytes / ms * 1e-6
        r = dict(kernel=name, mbytes=nbytes / 1e6, ms=ms, gbs=gbs, frac_of_hbm=gbs / hbm, note=note)
        print(json.dumps(r), flush=True)
        out.append(r)

    n = 64 * 1024 * 1024
    a, b = rnd(n), rnd(n)
    o = torch.empty_like(a)
    rec("torch_copy (reference point)", 16 * n, lambda: o.copy_(a))
    rec("k_ewise scale  (tt_scale core)", 16 * n, lambda: K.ewise(a, 0.5, out=o, rt=rt))
    rec("k_ewise axpy   (a + beta b)", 24 * n, lambda: K.ewise(a, 1.0, b, -1.0, out=o, rt=rt))
    rec("k_ewise sumsq  (norm, no store)", 8 * n, lambda: K.ewise(a, 1.0, want_sumsq=True, store=False, rt=rt))
    x4 = rnd(1024, 4, 4, 4096)
    sc = rnd(4).abs() + 0.5
    rec("k_permute4 (r,b,n,R)->(r,n,b,R) scaled", 16 * x4.numel(),
        lambda: K.permute4(x4, (0, 2, 1, 3), scale=sc, scale_axis=2, rt=rt))
    ca, cb = rnd(1024, 16, 1024), rnd(1024, 16, 1024)
    rec("k_block_diag mid   (tt_add middle core)", 8 * (ca.numel() + cb.numel() + 4 * ca.numel()),
        lambda: K.block_diag(ca, cb, "mid", rt=rt), "output 4x the inputs: zeros are written, not read")
    rec("k_block_diag first (tt_add first core)", 16 * (ca.numel() + cb.numel()), lambda: K.block_diag(ca, cb, "first", rt=rt))
    rec("k_block_diag last  (tt_add last core)", 16 * (ca.numel() + cb.numel()), lambda: K.block_diag(ca, cb, "last", rt=rt))
    e = rnd(4096, 2, 2, 2048)
    rec("k_embed IkronM (tt_IkronM core)", 8 * (e.numel() + 4 * e.numel()), lambda: K.embed(e, "IkronM", rt=rt))
    rec("k_embed MkronI (tt_MkronI core)", 8 * (e.numel() + 4 * e.numel()), lambda: K.embed(e, "MkronI", rt=rt))
    d = rnd(2048, 4, 2048)
    rec("k_embed diag   (tt_diag core)", 8 * (d.numel() + 4 * d.numel()), lambda: K.embed(d, "diag", rt=rt))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "bench_hbm.jsonl"), "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()

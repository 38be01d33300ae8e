"""Measures the two peaks BASELINE.md section 2 leaves open -- dense TF32 (tensor pipe) and FP32 SGEMM (FFMA pipe, TF32
off) -- with the method of MEASURED_PEAKS.json (torch.matmul 8192^3, best of 10 after warm-up, CUDA events), and writes
profiles/peaks_r2.json.  Run on the GPU box:  python tools/measure_peaks.py"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def best_tflops(a, b, reps=10):
    for _ in range(3):
        a @ b
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); a @ b; e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    n = a.shape[0]
    return 2.0 * n ** 3 / (best / 1e3) / 1e12


def main():
    dev = torch.device("cuda", 0)
    n = 8192
    a, b = torch.randn((n, n), device=dev), torch.randn((n, n), device=dev)
    out = {"how": "torch.matmul fp32 8192^3, best of 10, CUDA events; TF32 via torch.backends.cuda.matmul.allow_tf32",
           "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__}
    torch.backends.cuda.matmul.allow_tf32 = False
    out["fp32_sgemm_tflops"] = best_tflops(a, b)
    torch.backends.cuda.matmul.allow_tf32 = True
    out["tf32_tflops"] = best_tflops(a, b)
    torch.backends.cuda.matmul.allow_tf32 = False
    h = a.half(); g = b.half()
    out["fp16_tflops"] = best_tflops(h, g)
    path = os.path.join(ROOT, "gpurun_out" if "--scratch" in sys.argv else "profiles", "peaks_r2.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

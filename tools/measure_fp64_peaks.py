"""Measure the FP64 roofline denominators on the GPU box and write FP64_PEAKS.json (repo root).
MEASURED_PEAKS.json (driver-written) has HBM and bf16 only; SURVEY.md 8(d) asks the builder to add, in the
same style: (i) the DFMA peak (register-resident FMA chains), (ii) the DMMA peak via cuBLAS DGEMM 8192^3
(best of 10 and sustained), plus the raw DMMA.8x8x4 issue peak.  Run under gpurun:
    python tools/measure_fp64_peaks.py
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    out = {}
    exe = os.path.join(ROOT, "tools", "bin", "fp64_peaks")
    if not os.path.exists(exe):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe,
                               os.path.join(ROOT, "tools", "fp64_peaks.cu")])
    out["microbench"] = json.loads(subprocess.check_output([exe]).decode().strip().splitlines()[-1])
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    for _ in range(3):
        torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["cublas_dgemm_tflops"] = 2 * n ** 3 / (best * 1e-3) / 1e12
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cnt = 0
    t0 = time.time()
    while time.time() - t0 < 4.0:
        torch.matmul(a, b)
        cnt += 1
        if cnt % 4 == 0:
            torch.cuda.synchronize()
    e1.record()
    e1.synchronize()
    out["cublas_dgemm_tflops_sustained"] = 2 * n ** 3 * cnt / (e0.elapsed_time(e1) * 1e-3) / 1e12
    # symmetric rank-k update A^T A of a tall matrix (the Gram shape): 60000 x 2048
    A = torch.randn(60000, 2048, dtype=torch.float64, device="cuda")
    for _ in range(2):
        torch.matmul(A.t(), A)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(A.t(), A)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out["cublas_gram_60000x2048_ms"] = best
    out["cublas_gram_60000x2048_tflops_full"] = 2 * 60000 * 2048 * 2048 / (best * 1e-3) / 1e12
    out["gpu_name"] = torch.cuda.get_device_name(0)
    out["when"] = time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())
    out["how"] = ("tools/fp64_peaks.cu: 8 independent DFMA / DMMA.8x8x4 chains per thread, 148*8 CTAs x 256 threads, "
                  "best of 10 (burst) and back to back for 2 s (sustained); torch.matmul fp64 8192^3 best of 10 and 4 s loop")
    path = os.path.join(ROOT, "gpurun_out", "FP64_PEAKS.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    sys.exit(main())

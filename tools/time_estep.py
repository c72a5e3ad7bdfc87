"""Quick device timing of the K1 E-step at a given size (CUDA events on the library stream)."""
import ctypes
import json
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gingr_b200 import api, synthetic

M, N = int(sys.argv[1]), int(sys.argv[2])
ctx = api.Context(0)
fit = synthetic.fibonacci_sphere(M)
tgt = synthetic.make_target(synthetic.fibonacci_sphere(N), 0)
target = api.Target(ctx, tgt)
for _ in range(3):
    api.cpd_estep(ctx, target, fit, 25.0, 0.1)
stream = torch.cuda.ExternalStream(ctx.stream)
ts = []
import time
for _ in range(5):
    t0 = time.perf_counter()
    api.cpd_estep(ctx, target, fit, 25.0, 0.1)
    ts.append(time.perf_counter() - t0)
t = min(ts)
print(json.dumps({"M": M, "N": N, "estep_wall_ms": t * 1e3, "tflops_71": 71.0 * M * N / t / 1e12}))

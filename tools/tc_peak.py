"""Measured dense tensor-core peaks of this B200 for the roofline denominators bench.py uses next to
MEASURED_PEAKS.json (which has bf16 only): int8 (torch._int_mm -> cuBLASLt) and fp16 library GEMMs, 8192^3,
best of 10 (burst) and a 3-second back-to-back loop (sustained).  Writes profiles/measured_tc_peaks.json.

    python tools/tc_peak.py            (on the GPU box; a library GEMM is the yardstick here, not part of the product)
"""
import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def measure(fn, flops):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = max(best, flops / (a.elapsed_time(b) * 1e-3) / 1e12)
    n, t0 = 0, time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    while time.perf_counter() - t0 < 3.0:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    b.record(); torch.cuda.synchronize()
    return best, flops * n / (a.elapsed_time(b) * 1e-3) / 1e12


def main():
    n = 8192
    dev = torch.device("cuda", 0)
    out = {"gpu_name": torch.cuda.get_device_name(0), "how": "library GEMMs %d^3: best of 10 (burst) / 3 s loop (sustained)" % n}
    a8 = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    b8 = torch.randint(-8, 8, (n, n), dtype=torch.int8, device=dev)
    burst, sus = measure(lambda: torch._int_mm(a8, b8), 2.0 * n ** 3)
    out["int8_tops_burst"], out["int8_tops"] = burst, sus
    ah = torch.randn((n, n), dtype=torch.float16, device=dev)
    bh = torch.randn((n, n), dtype=torch.float16, device=dev)
    burst, sus = measure(lambda: torch.matmul(ah, bh), 2.0 * n ** 3)
    out["fp16_tflops_burst"], out["fp16_tflops"] = burst, sus
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    for path in (os.path.join(ROOT, "gpurun_out", "measured_tc_peaks.json"), os.path.join(ROOT, "profiles", "measured_tc_peaks.json")):
        json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

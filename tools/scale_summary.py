"""Collects the per-N bench lines of the round into profiles/scale_r02.json (SCALE format: one entry per GPU count,
whole-job value, max-over-ranks step time, efficiency against the single-GPU line of the same series).

    python tools/scale_summary.py            (reads profiles/bench_r02_*; run after copying the lines there)
"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SERIES = {
    "c3": ["bench_r02_c3_final.json", "bench_r02_c3_2gpu_final.json", "bench_r02_c3_4gpu_final.json", "bench_r02_c3_8gpu_final.json"],
    "c3_before_pipelining_1536_per_step": ["bench_r02_c3_v10.json", "bench_r02f_c3_2gpu.json", "bench_r02f_c3_4gpu.json", "bench_r02f_c3_8gpu.json"],
    "c4": ["bench_r02_c4_final.json", "bench_r02_c4_8gpu_final.json"],
    "c5": ["bench_r02_c5_final.json", "bench_r02_c5_2gpu_final.json", "bench_r02_c5_8gpu.json"],
}
out = {"format": "per-N lines of bench.py (weak scaling: per-GPU work fixed; value = whole-job fits/s, time = max over ranks)"}
for name, files in SERIES.items():
    rows = []
    for f in files:
        p = os.path.join(ROOT, "profiles", f)
        if not os.path.exists(p):
            continue
        d = json.loads(open(p).read().strip().splitlines()[-1])
        rows.append({"n_gpus": d["n_gpus"], "value": d["value"], "unit": d["unit"], "ms_per_step": d["ms_per_step"], "steps": d["steps"],
                     "warmup": d["warmup"], "replicates_per_gpu_per_step": d["config"]["replicates_per_gpu_per_step"],
                     "e2e": d["e2e"]["value"], "sm_mhz": d["clocks"]["sm_mhz"] if d.get("clocks") else None, "source": f})
    base = next((r for r in rows if r["n_gpus"] == 1), None)
    for r in rows:
        r["efficiency_vs_1gpu"] = r["value"] / (r["n_gpus"] * base["value"]) if base else None
    out[name] = rows
json.dump(out, open(os.path.join(ROOT, "profiles", "scale_r02.json"), "w"), indent=1)
for name, rows in out.items():
    if isinstance(rows, list):
        print(name, [(r["n_gpus"], round(r["value"]), round(r["efficiency_vs_1gpu"], 4) if r["efficiency_vs_1gpu"] else None) for r in rows])

"""Turns ncu reports (gpurun_out/*.ncu-rep, or the --page raw CSVs) into the tracked summaries under profiles/:

    python tools/ncu_summary.py <tag> <workload> <replicates_per_launch> stage=report.ncu-rep [stage=report ...]

  profiles/<stage>_<tag>_details.csv     metric,value of the captured launch (the metrics the judge greps)
  profiles/kernel_traffic.json           workload -> {stage: dram bytes (read + write) per launch, _replicates_per_launch}
bench.py scales the per-launch bytes to its replicate count; nothing is hand-typed."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__cluster_size", "sm__cycles_elapsed.max", "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed")
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def raw_rows(path):
    if path.endswith(".csv"):
        text = open(path).read()
    else:
        text = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    return list(csv.reader(io.StringIO(text)))


STAGE_OF = (("gram_mma_kernel", "gram_i8"), ("vote_mma_kernel", "cross"), ("gram_finalize_kernel", "finalize"), ("solve_kernel", "solve"),
            ("resample_images_kernel", "counts"), ("counts8_image_kernel", "colsum"), ("vote_c8_image_kernel", "colsum_vote"), ("counts_kernel", "counts"),
            ("gram_kernel", "gram"), ("reduce_chunks_kernel", "reduce"), ("scoregen_kernel", "scoregen"))


def num(v, u):
    return float(v.replace(",", "")) * UNIT.get(u, 1.0)


def main():
    """stage=report pairs name the stage of the LAST launch in each report; a bare report path holds several kernels and
    the last launch of each known kernel name is taken (STAGE_OF)."""
    tag, workload, reps = sys.argv[1], sys.argv[2], int(sys.argv[3])
    tp = os.path.join(ROOT, "profiles", "kernel_traffic.json")
    table = json.load(open(tp)) if os.path.exists(tp) else {}
    entry = {"_replicates_per_launch": reps, "_tag": tag}
    picked = {}
    for spec in sys.argv[4:]:
        stage, path = spec.split("=", 1) if "=" in spec else (None, spec)
        rows = raw_rows(path)
        hdr, units = rows[0], rows[1]
        kcol = hdr.index("Kernel Name")
        for vals in rows[2:]:
            st = stage
            if st is None:
                st = next((s_ for k_, s_ in STAGE_OF if vals[kcol].startswith(k_)), None)
            if st:
                picked[st] = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    for stage, rec in sorted(picked.items()):
        with open(os.path.join(ROOT, "profiles", "%s_%s_details.csv" % (stage, tag)), "w") as f:
            f.write("metric,unit,value\n")
            f.write("kernel,,%s\n" % rec["Kernel Name"][0].replace(",", ";"))
            for k in rec:
                if any(k == m or k.startswith(m) for m in KEEP) or ("stalled" in k and "per_warp_active" in k):
                    f.write("%s,%s,%s\n" % (k, rec[k][1], rec[k][0]))
        tot = sum(num(*rec[k]) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        entry[stage] = tot
        print("%-12s %-50s dram %.3e B/launch  %s %s" % (stage, rec["Kernel Name"][0][:50], tot, rec["gpu__time_duration.sum"][0], rec["gpu__time_duration.sum"][1]))
    if "colsum_vote" in entry:  # both image builders run under the bench stage "colsum"
        entry["colsum"] = entry.get("colsum", 0.0) + entry.pop("colsum_vote")
    table[workload] = entry
    json.dump(table, open(tp, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

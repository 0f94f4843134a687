"""Turn the ncu reports in gpurun_out/ into the small tracked files of profiles/ (run in the build container).

    python profiles/extract.py r1
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
NAMES = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
         'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
         'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed',
         'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
         'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
         'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
         'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
         'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'sass__inst_executed_local_loads']


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    res = []
    for r in rows[2:]:
        res.append({n: (r[hdr.index(n)], rows[1][hdr.index(n)]) for n in NAMES if n in hdr})
    return res


def to_bytes(v):
    val, unit = v
    return float(val.replace(",", "")) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[unit]


kernels = []
for rep in (f"prof_chamfer_{R}.ncu-rep", f"prof_stages_{R}.ncu-rep"):
    path = os.path.join(ROOT, "gpurun_out", rep)
    if os.path.exists(path):
        kernels += raw(path)

json.dump(kernels, open(os.path.join(ROOT, "profiles", f"ncu_{R}_metrics.json"), "w"), indent=1)
traffic = {}
for d in kernels:
    name = d["Kernel Name"][0]
    key = ("chamfer_nn_pair_kernel_n1000" if "nn_pair" in name
           else "chamfer_nn_walk_kernel_batch_8x32768" if "nn_walk_kernel<2, 8, 0" in name or "nn_walk_split_kernel<0" in name
           else "chamfer_nn_walk_kernel_24x32768" if "nn_walk" in name
           else "chamfer_nn_kernel_n1000" if "nn_kernel<8, 1, 0" in name or "nn_kernel<8, 1>" in name
           else "chamfer_nn_kernel_batch_sorted_8x32768" if "nn_kernel<4, 0, 1" in name
           else "chamfer_nn_kernel_merged_24x32768" if "nn_kernel" in name
           else "head_project_dusty1_b256_compact" if "head_project_image" in name
           else "head_project_dusty1_b256" if "head_project" in name
           else "chamfer_prep_sort_kd_24x32768" if "prep_sort_kernel<1" in name
           else "chamfer_prep_sort_24x32768" if "prep_sort" in name
           else "scan_preprocess_256x64x2048" if "scan_preprocess" in name
           else "fps_multi_888clouds" if "fps_multi" in name else "fps_pruned_148clouds")
    traffic[key] = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(traffic)            # keys captured in earlier rounds (other kernels, A/B paths) stay
json.dump(old, open(path, "w"), indent=1)

src = os.path.join(ROOT, "gpurun_out", f"launches_{R}.csv")
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    agg[r[ki]][0] += 1
    agg[r[ki]][1] += float(r[vi].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values())
with open(os.path.join(ROOT, "profiles", f"launches_{R}_summary.txt"), "w") as fh:
    for k, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1])[:12]:
        fh.write(f"{ms:12.3f} ms {n:5d}x {100 * ms / tot:6.2f}%  {k[:110]}\n")
import shutil
shutil.copy(src, os.path.join(ROOT, "profiles", f"launches_{R}.csv"))
for d in kernels:
    print({k: v[0][:50] for k, v in d.items() if k in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                                     "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
                                                     "smsp__issue_active.avg.pct_of_peak_sustained_active")})
print(open(os.path.join(ROOT, "profiles", f"launches_{R}_summary.txt")).read())

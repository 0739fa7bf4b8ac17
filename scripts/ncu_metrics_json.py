"""Extract the headline counters of every kernel in an .ncu-rep into a small JSON (profiles/*_kernel_metrics.json):
python scripts/ncu_metrics_json.py REPORT.ncu-rep OUT.json"""
import csv
import json
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__cluster_size",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",  # executed LDL / STL (spills that actually run)
           "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
           "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_global_ld.sum",
           "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv", "--metrics", ",".join(METRICS)], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
res = {}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "").replace("pft::", "")
    key = name
    k = 1
    while key in res:
        k += 1
        key = "%s#%d" % (name, k)
    res[key] = {"%s [%s]" % (h, u): v for h, u, v in zip(hdr, units, r) if "__" in h}
    res[key]["launch"] = "grid %s block %s" % (r[hdr.index("Grid Size")], r[hdr.index("Block Size")])
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
res["_source_hash"] = bench.source_hash()  # bench.py only reports the DRAM traffic of a capture taken on the sources it runs
json.dump(res, open(sys.argv[2], "w"), indent=1)
print("\n".join("%-40s %s us" % (k, v.get("gpu__time_duration.sum [us]")) for k, v in res.items() if isinstance(v, dict)))

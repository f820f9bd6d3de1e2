"""Turns an ncu report (ncu --set full ...) into the JSON summary committed under profiles/ (run here, no GPU needed):
    python tools/ncu_summary.py gpurun_out/r01_prof_c3b.ncu-rep profiles/r01_ncu_summary_c3b.json "<source note>"
One entry per kernel (first captured launch of each), holding the metrics the DESIGN.md tables and bench.py's
roofline.traffic are taken from."""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum",
    "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    # shared-memory data pipe: operand reads of the tensor core (tc) and everything else (lsu: TMA writes are not in it)
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__ops_path_tensor_op_utchmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
]


def main(rep, out, note):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    header, units = rows[0], rows[1]
    kernels = {}
    for r in rows[2:]:
        rec = dict(zip(header, r))
        name = rec["Kernel Name"]
        short = re.sub(r"^.*::", "", name.split("(")[0]).split("<")[0]
        if short in kernels:
            continue
        ent = {"kernel": name.split("(")[0], "grid": rec.get("Grid Size"), "block": rec.get("Block Size")}
        for m in KEEP:
            if m in rec and rec[m] != "":
                u = units[header.index(m)]
                try:
                    ent[f"{m} [{u}]"] = float(rec[m].replace(",", ""))
                except ValueError:
                    ent[f"{m} [{u}]"] = rec[m]
        kernels[short] = ent
    json.dump({"source": note, "kernels": kernels}, open(out, "w"), indent=1)
    for k, e in kernels.items():
        print(k, {x: e[x] for x in e if "time_duration" in x or "dram__bytes" in x})


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else "")

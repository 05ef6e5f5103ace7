"""Turn an .ncu-rep (brought back in gpurun_out/) into the compact summary kept under profiles/.
usage: python tests/tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r01_name.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__warps_eligible.avg.per_cycle_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red.sum", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_global_red.sum",
    "smsp__inst_executed_op_global_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of `{rep.split('/')[-1]}` (`ncu --set full --clock-control none`)", ""]
    seen = set()
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if name in seen:
            continue
        seen.add(name)
        lines += [f"## `{name}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                lines.append(f"| {k} | {r[i]} | {units[i]} |")
        stalls = [(hdr[i], r[i]) for i in range(len(hdr))
                  if "issue_stalled" in hdr[i] and hdr[i].endswith("per_issue_active.ratio")]

        def val(v):
            try:
                return float(v.replace(",", ""))
            except ValueError:
                return 0.0

        stalls.sort(key=lambda kv: -val(kv[1]))
        lines += ["", "top warp-stall reasons (warps stalled per issue-active cycle):", ""]
        for k, v in stalls[:7]:
            k = k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")
            lines.append(f"* {k}: {v}")
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

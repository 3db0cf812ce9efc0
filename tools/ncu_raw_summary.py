#!/usr/bin/env python
"""Key metrics of an `ncu --page raw --csv` export (one kernel per row): python tools/ncu_raw_summary.py file.csv [--md out.md]"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum"]
STALLS = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
STALL_NAMES = ["short_scoreboard", "long_scoreboard", "wait", "not_selected", "mio_throttle", "math_pipe_throttle", "dispatch_stall",
               "branch_resolving", "barrier", "lg_throttle", "no_instruction", "imc_miss", "drain", "membar", "sleeping", "tex_throttle", "selected"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = []
    for vals in rows[2:]:
        if len(vals) < len(hdr):
            continue
        name = vals[ix["Kernel Name"]].split("(")[0]
        out.append("## %s" % name)
        out.append("")
        out.append("| metric | value |")
        out.append("|---|---|")
        for w in WANT:
            if w in ix:
                out.append("| %s | %s %s |" % (w, vals[ix[w]], units[ix[w]]))
        st = []
        for s in STALL_NAMES:
            k = STALLS % s
            if k in ix and float(vals[ix[k]] or 0) >= 0.05:
                st.append("%s=%.2f" % (s, float(vals[ix[k]])))
        out.append("")
        out.append("Warp stall reasons (warps per issue-active cycle): " + ", ".join(st))
        out.append("")
    if "--json" in sys.argv:
        # per-kernel pipe utilisation for bench.py (roofline.pipes): keys = the library's profiler names
        import json
        alias = {"pesq_filter_tiled_kernel": "pesq_filter_kernel", "stoi_resample85_kernel": "stoi_resample_kernel"}
        keys = {"fp32_fma_pipe_pct": "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
                "alu_pipe_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
                "lsu_data_pipe_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                "warp_instructions": "smsp__inst_executed.sum"}
        pipes = {}
        for vals in rows[2:]:
            if len(vals) < len(hdr):
                continue
            name = vals[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
            name = alias.get(name, name)
            pipes[name] = {k: float(vals[ix[m]].replace(",", "")) for k, m in keys.items() if m in ix}
        json.dump({"_comment": "ncu --set full of one step at 8192 x 10 s (tools/gpu_evidence.sh); % of peak", "kernels": pipes},
                  open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
    txt = "\n".join(out)
    if "--md" in sys.argv:
        open(sys.argv[sys.argv.index("--md") + 1], "w").write(txt + "\n")
    print(txt)


if __name__ == "__main__":
    main()

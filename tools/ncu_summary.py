"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON/markdown for profiles/."""
import csv, io, json, subprocess, sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed.sum", "sm__inst_executed.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.avg.per_cycle_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum", "sm__sass_thread_inst_executed_op_fmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main(path, out_json=None):
    hdr, units, data = raw(path)
    res = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in KEYS or "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
                try:
                    d[h] = float(r[i].replace(",", ""))
                except ValueError:
                    d[h] = r[i]
                d[h + "__unit"] = units[i]
        res.append(d)
    if out_json:
        json.dump(res, open(out_json, "w"), indent=1)
    for d in res:
        print("==", d["kernel"][:90])
        for k in d:
            if k.endswith("__unit") or k == "kernel":
                continue
            v = d[k]
            if "stalled" in k and (not isinstance(v, float) or v < 2.0):
                continue
            print(f"  {k:95s} {v} {d.get(k + '__unit', '')}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)

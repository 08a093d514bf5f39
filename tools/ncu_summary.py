"""Summarise an `ncu --page raw --csv` / `--page source --csv` pair: key counters and the top stall sites."""
import csv
import sys

KEYS = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "l1tex__t_bytes.sum")


def main(raw, src=None, top=25):
    rows = list(csv.reader(open(raw)))
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        if h in KEYS:
            print("%-80s %-8s %s" % (h, u, v))
    if not src:
        return
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    print("total samples", tot, "instructions", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
    print("stall mix:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * c / max(tot, 1)) for h, c in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top]
    for i in order:
        r = data[i]
        n = int(r[ix["# Samples"]])
        st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print("%5d %7d %5.1f%% %-72s x%s %s" % (i, n, 100.0 * n / max(tot, 1), r[ix["Source"]].strip()[:72], r[ix["Instructions Executed"]], st))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None, int(sys.argv[3]) if len(sys.argv) > 3 else 25)

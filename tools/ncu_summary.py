"""Summarise an ncu report of one kernel launch: key raw metrics, stall reasons and the opcode mix with the share of
warp-stall samples per opcode (needs -lineinfo / --import-source on for the source page).
Usage: python tools/ncu_summary.py report.ncu-rep"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum ", "dram__bytes_write.sum ", "launch__registers_per_thread ",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum ", "sm__ops_path_tensor_src_fp64.sum ", "lts__t_bytes.sum ", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum "]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("## raw metrics")
    for h, u, v in zip(hdr, units, vals):
        if any((h + " ").startswith(k) for k in KEYS):
            print("%-80s %12s %s" % (h, v, u))
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    byop, ex, stalls = collections.Counter(), collections.Counter(), collections.Counter()
    per = collections.defaultdict(collections.Counter)
    for r in data:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        op = m.group(2).split(".")[0] if m else "?"
        byop[op] += int(r[ix["# Samples"]])
        ex[op] += int(r[ix["Instructions Executed"]])
        for c in stall_cols:
            v = int(r[ix[c]] or 0)
            stalls[c] += v
            per[op][c] += v
    print("## warp-stall samples: %d" % tot)
    print("  " + ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / tot) for c, v in stalls.most_common(9)))
    print("## opcode: share of samples | share of executed instructions | top stalls")
    te = sum(ex.values())
    for op, n in byop.most_common(12):
        print("  %-8s %6.2f%% | %5.1f%% | %s" % (op, 100.0 * n / tot, 100.0 * ex[op] / te,
                                                ", ".join("%s %.1f%%" % (c[6:], 100.0 * v / tot) for c, v in per[op].most_common(3))))


if __name__ == "__main__":
    main(sys.argv[1])

"""Summarise an ncu report (.ncu-rep) on a box without a GPU: headline metrics of the first kernel in
the report and the source lines where the warp-stall samples concentrate.

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [n_lines] [kernel-name-regex] > profiles/rNN_<kernel>_ncu_summary.txt
"""
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
]


KERNEL = []


def run(args):
    return subprocess.run(["ncu", "-i"] + args + KERNEL, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    n_lines = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    if len(sys.argv) > 3:
        KERNEL.extend(["--kernel-name", "regex:" + sys.argv[3]])
    raw = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    if len(raw) >= 3:
        names, units, vals = raw[0], raw[1], raw[2]
        col = {n: i for i, n in enumerate(names)}
        print("kernel:", vals[col.get("Kernel Name", 4)])
        for m in RAW:
            if m in col:
                print("%-64s %s %s" % (m, vals[col[m]], units[col[m]]))
    # per-line stall samples: "cuda,sass" view, one block per source file; rows with a line number and
    # "-" as address are the per-source-line aggregates
    out = run([rep, "--page", "source", "--csv", "--print-source", "cuda,sass"])
    cur, header = None, None
    rows = []
    for row in csv.reader(io.StringIO(out)):
        if not row:
            continue
        if row[0] == "File Path":
            cur, header = row[1], None
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            header = row
            continue
        if header and cur and len(row) == len(header) and row[2] == "-":
            d = {}
            for k, v in zip(header[4:], row[4:]):
                d[k] = v
            try:
                samp = int(d.get("# Samples", "0") or 0)
                inst = int(d.get("Instructions Executed", "0") or 0)
            except ValueError:
                continue
            if samp or inst:
                rows.append((samp, inst, cur.split("/")[-1], row[0], row[1].strip(), d, cur))
    tot_s = sum(r[0] for r in rows) or 1
    tot_i = sum(r[1] for r in rows) or 1
    print("total warp-instructions %d, stall samples %d" % (tot_i, tot_s))
    stall_cols = [k for k in (rows[0][5].keys() if rows else []) if k.startswith("stall_") and "Not Issued" not in k]
    totals = {k: 0 for k in stall_cols}
    for r in rows:
        for k in stall_cols:
            try:
                totals[k] += int(r[5][k] or 0)
            except ValueError:
                pass
    print("stall totals:", sorted(((v, k) for k, v in totals.items() if v), reverse=True)[:10])
    # per-function aggregates: a source line belongs to the last function header above it
    import os
    import re
    fn_of = {}
    for path in sorted({r[6] for r in rows}):
        if not os.path.exists(path):
            continue
        cur_fn, names = "(file scope)", {}
        pending = False
        for i, line in enumerate(open(path, errors="replace"), 1):
            if re.match(r"^(template\s*<|__device__|__global__|static\s+__device__)", line):
                pending = True
            if pending:
                m = re.search(r"([A-Za-z_][A-Za-z_0-9]*)\s*\(", line)
                if m and not line.startswith("template"):
                    cur_fn, pending = m.group(1), False
            names[i] = cur_fn
        fn_of[path] = names
    agg = {}
    for samp, inst, f, ln, src, d, path in rows:
        fn = fn_of.get(path, {}).get(int(ln), f)
        a = agg.setdefault(fn, [0, 0, {}])
        a[0] += inst
        a[1] += samp
        for k in stall_cols:
            try:
                a[2][k] = a[2].get(k, 0) + int(d[k] or 0)
            except ValueError:
                pass
    print("--- by function (inlined code is attributed to the function its source line is in)")
    for fn, (inst, samp, st) in sorted(agg.items(), key=lambda x: -x[1][0]):
        top = ", ".join("%s %d" % (k[6:], v) for k, v in sorted(st.items(), key=lambda x: -x[1])[:4] if v)
        print("inst %5.1f%% samp %5.1f%%  %-26s %s" % (100.0 * inst / tot_i, 100.0 * samp / tot_s, fn, top))
    print("--- by samples")
    for samp, inst, f, ln, src, d, _ in sorted(rows, key=lambda r: -r[0])[:n_lines]:
        top = sorted(((int(d[k] or 0), k) for k in stall_cols), reverse=True)[:2]
        print("samp %5.1f%% inst %5.1f%%  %s:%s  %s  %s" % (100.0 * samp / tot_s, 100.0 * inst / tot_i, f, ln, src[:90], top))
    print("--- by instructions")
    for samp, inst, f, ln, src, d, _ in sorted(rows, key=lambda r: -r[1])[:n_lines]:
        print("inst %5.1f%% samp %5.1f%%  %s:%s  %s" % (100.0 * inst / tot_i, 100.0 * samp / tot_s, f, ln, src[:90]))


if __name__ == "__main__":
    main()

"""tools/ncu_csv_summary.py -- markdown table of the tracked metrics from `ncu --page raw --csv` exports
(the .ncu-rep files are turned into CSV on the GPU box so that gpurun_out stays small).
usage: python tools/ncu_csv_summary.py file.raw.csv [...]"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "time us", 1e-3), ("launch__grid_size", "grid", 1), ("launch__block_size", "block", 1),
        ("launch__registers_per_thread", "regs", 1), ("dram__bytes_read.sum", "DRAM rd MB", None), ("dram__bytes_write.sum", "DRAM wr MB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe %", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1),
        ("sm__inst_executed_pipe_fp64.sum", "FP64 inst", 1),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts", 1),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts", 1),
        ("smsp__inst_executed.sum", "warp inst", 1)]


def units(hdr_units, name, val):
    return val


def main():
    for path in sys.argv[1:]:
        r = list(csv.reader(open(path)))
        hdr, unit, rows = r[0], r[1], r[2:]
        col = {h: i for i, h in enumerate(hdr)}
        print("### %s\n" % path.split("/")[-1])
        names = [k for k, _, _ in WANT if k in col]
        print("| kernel | " + " | ".join(lbl for k, lbl, _ in WANT if k in col) + " | top stalls |")
        print("|---|" + "---|" * (len(names) + 1))
        stall_cols = [(h, i) for i, h in enumerate(hdr) if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and "not_issued" not in h]
        for row in rows:
            kn = row[col["Kernel Name"]]
            kn = kn.split("(")[0].replace("void ", "")
            cells = []
            for k, lbl, sc in WANT:
                if k not in col:
                    continue
                v = row[col[k]].replace(",", "")
                u = unit[col[k]]
                try:
                    f = float(v)
                except ValueError:
                    cells.append(v); continue
                if k == "gpu__time_duration.sum":
                    f = f * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                    cells.append("%.1f" % f)
                elif sc is None:
                    f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
                    cells.append("%.1f" % f)
                elif f >= 1e6:
                    cells.append("%.3g" % f)
                else:
                    cells.append(("%.1f" % f) if f != int(f) else "%d" % int(f))
            st = sorted(((float(row[i] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h, i in stall_cols), reverse=True)[:3]
            print("| %s | %s | %s |" % (kn, " | ".join(cells), ", ".join("%s %.1f" % (h, v) for v, h in st)))
        print()


if __name__ == "__main__":
    main()

"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
    python tools/summarize_ncu.py launches <launches.csv> <out.md> [title]
    python tools/summarize_ncu.py full <report.ncu-rep> <out.md> [kernel regex]
"""
import collections, csv, io, re, subprocess, sys


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] == "ns" else (v * 1000 if r[ui] == "ms" else v)
        agg.setdefault(r[ki], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` launch list "
                "(cold-cache, serialised: compare shares, not absolutes).\n\n")
        f.write("| kernel | launches | mean us | min us | total us | share |\n|---|---:|---:|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            name = re.sub(r"\(.*", "", k)[:90]
            f.write(f"| `{name}` | {len(v)} | {sum(v)/len(v):.2f} | {min(v):.2f} | {sum(v):.1f} | {100*sum(v)/tot:.1f}% |\n")
    print("wrote", out)


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__grid_size",
        "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tc.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def full(path, out, regex):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    with open(out, "w") as f:
        f.write(f"# ncu --set full: {path.split('/')[-1]}\n\n")
        for r in rows[2:]:
            if regex and not re.search(regex, r[ki]):
                continue
            f.write(f"## `{r[ki][:100]}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in hdr and r[hdr.index(w)] != "":
                    f.write(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |\n")
            tensor = [(h, r[i]) for i, h in enumerate(hdr) if ("tensor" in h or "pipe_tc" in h) and r[i] not in ("", "0")]
            for h, v in tensor[:8]:
                f.write(f"| {h} | {v} | |\n")
            f.write("\n")
        det = subprocess.run(["ncu", "-i", path, "--page", "details"], capture_output=True, text=True).stdout
        keep = [l for l in det.splitlines() if re.search(r"Duration|Memory Throughput|DRAM Throughput|L2 Hit|L1/TEX Hit|Eligible|Issued Ipc|"
                                                         r"Warp Cycles Per Issued|Achieved Occupancy|Theoretical Occupancy|Registers Per|stall|"
                                                         r"highest-utilized|Waves Per SM|^  [a-zA-Z].*\(", l)]
        f.write("## details page (selected lines)\n\n```\n" + "\n".join(keep[:80]) + "\n```\n")
    print("wrote", out)


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "launch list")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else None)

#!/usr/bin/env python
"""Turn ncu output into the text summaries kept under profiles/.

    python tools/summarize_ncu.py launches gpurun_out/launches.csv            # per-kernel time shares
    python tools/summarize_ncu.py report   gpurun_out/prof.ncu-rep [min_exec] # key metrics + stall profile
"""
import collections, csv, io, re, subprocess, sys


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("=="))]
    h = rows[0]
    iname, imetric, ival = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= ival or r[imetric] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[iname].replace("rp::<unnamed>::", "").replace("void ", ""))[:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[ival].replace(",", "")) / 1e3  # ns -> us
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':70s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:70s} {n:8d} {us:10.1f} {100 * us / tot:6.1f}%")
    print(f"{'total':70s} {sum(v[0] for v in agg.values()):8d} {tot:10.1f}")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct"]


def report(path, min_exec=0):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]
    d = dict(zip(h, zip(u, v)))
    print("kernel:", d.get("Kernel Name", ("", "?"))[1])
    for k in KEYS:
        if k in d:
            print(f"  {k:82s} {d[k][1]:>14s} {d[k][0]}")
    for k, (uu, vv) in d.items():
        if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio") and float(vv or 0) >= 0.05:
            print(f"  {k:82s} {float(vv):14.3f}")
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h, data = rows[1], rows[2:]
    isrc, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
    cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    hot = [r for r in data if int(r[iex] or 0) >= min_exec]
    n = sum(int(r[isamp] or 0) for r in hot) or 1
    agg, byop = collections.Counter(), collections.Counter()
    for r in hot:
        for i in cols:
            if r[i]:
                agg[h[i][6:]] += int(r[i])
        op = r[isrc].split()[1 if r[isrc].startswith("@") else 0].split(".")[0]
        byop[op] += int(r[isamp] or 0)
    print(f"warp-state samples over {len(hot)} instructions executed >= {min_exec} times: {n}")
    print("  by stall reason:", ", ".join(f"{k} {100 * c / n:.1f}%" for k, c in agg.most_common(10)))
    print("  by opcode:      ", ", ".join(f"{k} {100 * c / n:.1f}%" for k, c in byop.most_common(12)))
    print("  top instructions:")
    for r in sorted(hot, key=lambda r: -int(r[isamp] or 0))[:12]:
        st = sorted(((h[i][6:], int(r[i])) for i in cols if r[i] and int(r[i]) > 0), key=lambda kv: -kv[1])[:2]
        print(f"    {int(r[isamp]):6d}  {r[isrc][:64]:64s} {st}")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)

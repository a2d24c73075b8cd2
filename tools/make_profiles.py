"""Turn ncu exports in gpurun_out/ into the committed summaries under profiles/.

    python tools/make_profiles.py <raw.csv from `ncu --page raw --csv`> <launches.csv> <tag>
"""
import collections, csv, json, sys
raw, launches, tag = sys.argv[1], sys.argv[2], sys.argv[3]
# `raw` may be a comma-separated list of exports (one per capture); columns differ between captures
captures = []
for path in raw.split(","):
    rows = list(csv.reader(open(path)))
    if len(rows) > 2:
        captures.append((rows[0], rows[1], rows[2:]))
M = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % peak"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
     ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU wavefronts % peak"),
     ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % peak"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
     ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs/thread"),
     ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp instructions")]
seen = collections.OrderedDict()
for hdr, units, data in captures:
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        name = r[idx["Kernel Name"]]
        seen.setdefault(name, []).append((r, idx, units))
out = [f"# ncu --set full summary ({tag})", "",
       "Captured with `ncu --set full --clock-control none --import-source on` on one B200 (one launch per",
       "row: the LAST captured launch of each kernel, i.e. after warm-up launches of the same process).",
       "Times under ncu are serialised and cold-cache; bench.py's CUDA-event numbers are the ones reported.", ""]
traffic = {}
for name, rs in seen.items():
    r, idx, units = rs[-1]
    short = name.replace("void b200::<unnamed>::", "").replace("b200::<unnamed>::", "")[:110]
    out += [f"## `{short}`", "", "| metric | value |", "|---|---|"]
    for m, label in M:
        if m in idx:
            v = r[idx[m]]
            try:
                v = f"{float(v):,.3f}".rstrip("0").rstrip(".")
            except ValueError:
                pass
            out.append(f"| {label} | {v} {units[idx[m]]} |")
    def f(m):
        v = float(r[idx[m]]); u = units[idx[m]]
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
    traffic[short.split("(")[0]] = f("dram__bytes_read.sum") + f("dram__bytes_write.sum")
    out.append("")
open(f"profiles/{tag}_ncu_hot_kernels.md", "w").write("\n".join(out))
# launch list
rows = [r for r in csv.reader(open(launches)) if len(r) > 10]
h = rows[0]; ix = {k: i for i, k in enumerate(h)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]]) / (1e3 if r[ix["Metric Unit"]] == "ns" else 1)
    a = agg.setdefault(r[ix["Kernel Name"]], [0, 0.0, r[ix["Block Size"]], r[ix["Grid Size"]]])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
L = [f"# Launch list of `python bench.py --steps 2 --warmup 3 --no-cpu` ({tag})", "",
     "`ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv`; per-launch times are cold-cache and",
     "serialised — compare SHARES.  torch kernels (random number generation, sort of the query batches, copies) set",
     "up the synthetic inputs outside the timed regions.", "", "| launches | total ms | share | block | grid (last) | kernel |", "|---|---|---|---|---|---|"]
for k, (n, t, b, g) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    L.append(f"| {n} | {t/1e3:.3f} | {100*t/tot:.1f}% | {b} | {g} | `{k[:120]}` |")
open(f"profiles/{tag}_launch_list.md", "w").write("\n".join(L) + "\n")
import shutil
shutil.copy(launches, f"profiles/{tag}_launches.csv")
print(json.dumps(traffic, indent=1))
# roofline_traffic.json: the figure bench.py reports as roofline.traffic, stamped with a hash of the kernel sources it
# was captured for (bench.py flags it STALE when the sources have changed since)
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
key = [k for k in traffic if "interp2_scattered_smem_kernel<double, 2>" in k]
if key:
    json.dump({"interp2_scattered_f64": {"dram_bytes_per_launch": int(traffic[key[0]]),
                                         "source_sha1": bench.source_stamp(bench.INTERP2_SOURCES), "capture": tag,
                                         "source": f"profiles/{tag}_ncu_hot_kernels.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch, 1e8 queries, tile layout)"}},
              open("profiles/roofline_traffic.json", "w"), indent=1)

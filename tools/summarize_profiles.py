"""Turn the ncu outputs in gpurun_out/ into the tracked summaries under profiles/.

    python tools/summarize_profiles.py <launch_list.csv> <capture.ncu-rep | capture_raw.csv>[,<more>...] <tag>

Writes profiles/<tag>_launches.md (per-kernel totals and shares of the step), profiles/<tag>_kernels.md
(key metrics of every captured launch) and profiles/ncu_traffic.json (DRAM bytes per launch of
the hot kernels, read back by bench.py for roofline.traffic)."""
import collections, csv, json, re, subprocess, sys
from pathlib import Path

REPO = Path(__file__).resolve().parents[1]
launch_csv, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]

def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.split("::")[-1]
    return name.replace("void ", "").strip()

# ---- launch list
rows = list(csv.DictReader([l for l in open(launch_csv) if not l.startswith("==")]))
seq = []
for r in rows:
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    us = v / 1000 if u in ("ns", "nsecond") else (v * 1000 if u in ("ms", "msecond") else v)
    seq.append((short(r["Kernel Name"]), us))
agg = collections.OrderedDict()
for n, us in seq:
    a = agg.setdefault(n, [0, 0.0, 1e30, 0.0]); a[0] += 1; a[1] += us; a[2] = min(a[2], us); a[3] = max(a[3], us)
# steady-state step: median duration of the last launches of each loop kernel (the timed region of
# bench.py ends the capture; medians ignore the few first-iteration / re-run launches among them)
import statistics
loop = [(n, us) for n, us in seq if n.startswith(("k_match_grid", "k_reject_solve", "k_rs_fused", "k_bf_"))]
last = collections.defaultdict(list)
for n, us in loop[-120:]:
    last[n].append(us)
step_us = {n: statistics.median(v) for n, v in last.items() if len(v) >= 5}
tot = sum(step_us.values())
out = [f"# {tag}: ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`, cold caches, serialised)\n",
       "Command: `python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 1 --no-c4c5` under ncu (first 600 launches); absolute times are",
       "cold-cache and serialised, the SHARES are what is comparable with the CUDA-event numbers of bench.py.\n",
       "## All launches by kernel\n", "| kernel | launches | total us | min us | max us |", "|---|---:|---:|---:|---:|"]
for n, (c, t, mn, mx) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"| `{n}` | {c} | {t:.1f} | {mn:.1f} | {mx:.1f} |")
out += ["\n## Steady-state iteration (last launches of the timed region)\n", "| kernel | median us per launch | share |", "|---|---:|---:|"]
for n, t in sorted(step_us.items(), key=lambda kv: -kv[1]):
    out.append(f"| `{n}` | {t:.1f} | {100 * t / tot:.1f} % |")
(REPO / "profiles" / f"{tag}_launches.md").write_text("\n".join(out) + "\n")

# ---- full capture(s)
want = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("lts__t_bytes.sum", "L2 bytes"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("smsp__inst_executed.sum", "warp inst"), ("smsp__inst_executed_pipe_fma.sum", "FMA-pipe inst"),
        ("smsp__inst_executed_pipe_fp64.sum", "FP64-pipe inst"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %")]
def num(x):
    try: return float(x.replace(",", ""))
    except ValueError: return float("nan")
def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
lines = [f"# {tag}: ncu --set full captures (`--clock-control none --import-source on`)\n",
         "| kernel | " + " | ".join(w[1] for w in want) + " |", "|---|" + "---:|" * len(want)]
traffic = collections.defaultdict(list)
rows_all = []
for one in rep.split(","):
    # a .ncu-rep (converted here) or the raw CSV page already written on the GPU box
    raw = (open(one).read() if one.endswith(".csv") else
           subprocess.run(["ncu", "-i", one, "--page", "raw", "--csv"], capture_output=True, text=True).stdout)
    r = list(csv.reader([ln for ln in raw.splitlines() if ln.startswith('"')]))
    hdr, units, data = r[0], r[1], r[2:]
    rows_all.append(({h: i for i, h in enumerate(hdr)}, units, data))
for idx, units, data in rows_all:
  for d in data:
      name = short(d[idx["Kernel Name"]])
      cells = []
      for m, _ in want:
          if m in idx:
              v, u = num(d[idx[m]]), units[idx[m]]
              cells.append(f"{v:,.1f} {u}" if u not in ("",) else f"{v:,.0f}")
          else:
              cells.append("-")
      lines.append(f"| `{name}` | " + " | ".join(cells) + " |")
      rd = to_bytes(num(d[idx["dram__bytes_read.sum"]]), units[idx["dram__bytes_read.sum"]])
      wr = to_bytes(num(d[idx["dram__bytes_write.sum"]]), units[idx["dram__bytes_write.sum"]])
      traffic[name.split("<")[0]].append(rd + wr)
(REPO / "profiles" / f"{tag}_kernels.md").write_text("\n".join(lines) + "\n")
tj = {k: float(sorted(v)[len(v) // 2]) for k, v in traffic.items()}
tj["_note"] = "median dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full (caches flushed before each launch)"
(REPO / "profiles" / "ncu_traffic.json").write_text(json.dumps(tj, indent=1) + "\n")
print("\n".join(out[-8:])); print(json.dumps(tj))

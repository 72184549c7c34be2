#!/usr/bin/env python
"""Summarise the round's ncu evidence from gpurun_out/ into profiles/ (README.md table, ncu_traffic.json, launch list).
usage: python tools/prof_summary.py r01c"""
import csv, io, json, os, re, subprocess, sys, collections, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r02"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles")
KIND = {"0": "lock_2pl", "1": "lock_fasst", "2": "log", "3": "store", "4": "tatp", "5": "smallbank"}

def short(name):
    m = re.search(r"(k_\w+)<(\d)", name)
    return f"{m.group(1)}<{KIND.get(m.group(2), m.group(2))}>" if m else re.sub(r"\(.*", "", name)[:60]

def launches(path):
    agg = collections.OrderedDict()
    lines = [l for l in open(path) if not l.startswith("==")]
    for row in csv.DictReader(lines):
        try: v = float(row["Metric Value"].replace(",", ""))
        except Exception: continue
        u = row["Metric Unit"]
        v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        a = agg.setdefault(short(row["Kernel Name"]), [0, 0.0]); a[0] += 1; a[1] += v
    return agg

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    return [dict(zip(hdr, zip(r, units))) for r in data]

def num(cell):
    v, u = cell
    v = float(v.replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "ns": 1e-3, "us": 1, "ms": 1e3}
    return v * scale.get(u, 1)

traffic = {}
md = [f"# profiles/ -- capture `{R}` (B200, `--clock-control none --cache-control none`)", "",
      "Produced by `tools/round_capture.sh` under `gpurun`, summarised by `tools/prof_summary.py`:", "```"]
md += [l.rstrip() for l in open(os.path.join(ROOT, "tools", "round_capture.sh")) if "ncu --" in l or l.startswith("python bench")]
md += ["```", ""]
lp = os.path.join(G, f"launches_{R}.csv")
if os.path.exists(lp):
    shutil.copy(lp, os.path.join(P, f"{R}_launches_bench_steps2.csv"))
    agg = launches(lp); tot = sum(a[1] for a in agg.values())
    md += [f"## Launch list of the bench command (`{R}_launches_bench_steps2.csv`; cold-cache, serialised: compare SHARES)", "",
           "| kernel | launches | total us | avg us | share |", "|---|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    md.append("")
md += ["## `--set full` captures (one chunk = 1,048,576 requests per launch)", "",
       "| workload | kernel | duration us | DRAM read MB | DRAM write MB | DRAM % of peak | L2 hit % | achieved occupancy % | regs | grid x block |", "|---|---|---|---|---|---|---|---|---|---|"]
for wl in ("fasst", "store"):
    rep = os.path.join(G, f"prof_{wl}_{R}.ncu-rep")
    if not os.path.exists(rep): continue
    for r in raw(rep):
        name = short(r["Kernel Name"][0])
        rd, wr = num(r["dram__bytes_read.sum"]), num(r["dram__bytes_write.sum"])
        md.append("| %s | `%s` | %.1f | %.1f | %.1f | %s | %s | %s | %s | %s x %s |" % (
            wl, name, num(r["gpu__time_duration.sum"]), rd / 1e6, wr / 1e6,
            r.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ("?", ""))[0],
            r.get("lts__t_sector_hit_rate.pct", ("?", ""))[0],
            r.get("sm__warps_active.avg.pct_of_peak_sustained_active", ("?", ""))[0],
            r.get("launch__registers_per_thread", ("?", ""))[0], r.get("launch__grid_size", ("?", ""))[0], r.get("launch__block_size", ("?", ""))[0]))
# steady-state traffic: the single-pass captures (no kernel replay), last launch of each kernel
md += ["", "## Steady-state DRAM traffic per launch (ONE metrics pass per kernel, no replay: the L2 holds what the preceding launches left)", "",
       "| capture | kernel | us | DRAM read MB | DRAM write MB | L2 read / write / atom / red Msectors |", "|---|---|---|---|---|---|"]
for wl in ("fasst", "store", "route"):
    cp = os.path.join(G, f"dram_{wl}_{R}.csv")
    if not os.path.exists(cp): continue
    shutil.copy(cp, os.path.join(P, f"{R}_dram_{wl}.csv"))
    rows = [r for r in csv.reader(l for l in open(cp) if not l.startswith("==")) if len(r) > 10]
    h = rows[0]; iK, iM, iV, iID = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        per.setdefault((r[iID], short(r[iK])), {})[r[iM]] = float(r[iV].replace(",", ""))
    for (i, name), m in per.items():
        rd, wr = m.get("dram__bytes_read.sum", 0), m.get("dram__bytes_write.sum", 0)
        traffic[name] = int(rd + wr)                       # the last launch of a kernel wins: the steadiest one
        md.append("| %s #%s | `%s` | %.1f | %.1f | %.1f | %.2f / %.2f / %.2f / %.2f |" % (wl, i, name, m.get("gpu__time_duration.sum", 0) / 1e3, rd / 1e6, wr / 1e6,
                  m.get("lts__t_sectors_op_read.sum", 0) / 1e6, m.get("lts__t_sectors_op_write.sum", 0) / 1e6, m.get("lts__t_sectors_op_atom.sum", 0) / 1e6, m.get("lts__t_sectors_op_red.sum", 0) / 1e6))
traffic["_source"] = f"dram__bytes_read.sum + dram__bytes_write.sum per launch, ONE metrics pass per kernel (no replay), --cache-control none --clock-control none: profiles/{R}_dram_*.csv (steady state, 2^20 requests per launch)"
md += ["", "`ncu_traffic.json` = those sums, read by bench.py for `roofline.traffic` (a STATIC figure: bench.py says so).", ""]
sp = os.path.join(G, f"sass_{R}.txt")
if os.path.exists(sp):
    shutil.copy(sp, os.path.join(P, f"{R}_sass_excerpt.txt"))
    md += [f"`{R}_sass_excerpt.txt`: per kernel, how many `UBLKCP` (TMA bulk copies), `SYNCS` (mbarrier), `MATCH`, `ATOMG` / `RED` instructions the built library holds (`cuobjdump -sass`).", ""]
bj = os.path.join(G, f"bench_{R}_n1.json")
if os.path.exists(bj):
    shutil.copy(bj, os.path.join(P, f"{R}_bench_n1.json"))
    md += [f"The headline JSON line of the same build (plain run, not under ncu) is `{R}_bench_n1.json`.", ""]
# bench.py keys: k_apply<lock_fasst>, k_apply<store>
json.dump(traffic, open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
open(os.path.join(P, f"{R}_README.md"), "w").write("\n".join(md))
print("\n".join(md)); print(traffic)

"""Trim an `ncu -i X.ncu-rep --page raw --csv` dump to the columns quoted in profiles/ and print a markdown table
(per kernel + grid: the launch with the median duration).  usage: python tools/ncu_trim.py raw.csv trimmed.csv"""
import csv, re, sys, statistics

COLS = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
rows = list(csv.reader(open(sys.argv[1], newline="")))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = [hdr.index(c) if c in hdr else -1 for c in COLS]
unit = {c: (units[i] if i >= 0 else "") for c, i in zip(COLS, idx)}


def num(v):
    try:
        return float(v.replace(",", ""))
    except Exception:
        return float("nan")


def to_us(v, u):
    return v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)


def to_mb(v, u):
    return {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * v


out = []
for r in data:
    if len(r) < len(hdr):
        continue
    out.append([r[i] if i >= 0 else "" for i in idx])
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(COLS)
    w.writerow([unit[c] for c in COLS])
    w.writerows(out)
groups = {}
for r in out:
    name = re.sub(r"\(.*", "", r[0]).replace("fedb200::", "").replace("void ", "")
    groups.setdefault((name, r[1]), []).append(r)
print("| kernel | grid | launches | µs | regs | SM throughput % | tensor pipe active % | DRAM throughput % | DRAM read / write MB | L2 throughput % | L2 hit % | L2→SM MB | warps active % |")
print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
for (name, grid), rs in sorted(groups.items()):
    rs.sort(key=lambda r: num(r[3]))
    m = rs[len(rs) // 2]
    print("| `%s` | %s | %d | %.1f | %s | %.1f | %.1f | %.1f | %.1f / %.1f | %.1f | %.1f | %.1f | %.1f |" % (
        name, grid.replace(", 1, 1", ""), len(rs), to_us(num(m[3]), unit[COLS[3]]), m[4], num(m[5]), num(m[6]), num(m[7]),
        to_mb(num(m[8]), unit[COLS[8]]), to_mb(num(m[9]), unit[COLS[9]]), num(m[10]), num(m[11]), to_mb(num(m[12]), unit[COLS[12]]), num(m[13])))

"""Per-kernel table from an `ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --csv` launch list:
the launches of ONE step (the last complete run of the kernel sequence before the cut-off), serialised and cold.
    python profiles/launch_table.py gpurun_out/launches.csv [first_kernel_of_a_step]"""
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10 and r[0].isdigit()]
first = sys.argv[2] if len(sys.argv) > 2 else "k_edjob"
which = int(sys.argv[3]) if len(sys.argv) > 3 else -2          # which run of the kernel sequence (default: the last complete one)
by_id = {}
for r in rows:
    kid, name, metric, val = r[0], r[4].split("(")[0], r[-3], r[-1]
    by_id.setdefault(kid, {"name": name})[metric] = float(val.replace(",", ""))
launches = [by_id[k] for k in sorted(by_id, key=int)]
starts = [i for i, l in enumerate(launches) if l["name"] == first]
if len(starts) >= 2:
    a = starts[which]
    b = [x for x in starts if x > a]
    launches = launches[a:(b[0] if b else len(launches))]
tot = 0.0
print("%-16s %10s %14s" % ("kernel", "ms", "warp instr"))
for l in launches:
    ms = l.get("gpu__time_duration.sum", 0.0) / 1e6
    tot += ms
    print("%-16s %10.3f %14.0f" % (l["name"], ms, l.get("smsp__inst_executed.sum", 0)))
print("%-16s %10.3f" % ("sum", tot))

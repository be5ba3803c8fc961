"""dram traffic per kernel launch from an ncu launch list (`--metrics gpu__time_duration.sum,smsp__inst_executed.sum,
dram__bytes_read.sum,dram__bytes_write.sum --csv`): the launches of ONE batch (between two k_smtjob launches of the
full grid), largest launch per kernel name (the early walk of hz_early_hash repeats some kernels on one slot).
    python profiles/launch_traffic.py gpurun_out/launches.csv rollup-main-2048 > profiles/r2_dram_traffic.json"""
import csv
import json
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 10 and r[0].isdigit()]
by_id = {}
for r in rows:
    kid, name, metric, unit, val = r[0], r[4].split("(")[0], r[-3], r[-2], r[-1]
    v = float(val.replace(",", ""))
    if metric.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    by_id.setdefault(int(kid), {"name": name})[metric] = v
launches = [by_id[k] for k in sorted(by_id)]
per, ms, inst = {}, {}, {}
for l in launches:
    t = l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    if t >= per.get(l["name"], -1.0):
        per[l["name"]] = t
        ms[l["name"]] = l.get("gpu__time_duration.sum", 0.0) / 1e6
        inst[l["name"]] = l.get("smsp__inst_executed.sum", 0.0)
print(json.dumps({sys.argv[2]: {k: int(v) for k, v in per.items()},
                  "ms_under_ncu": {k: round(v, 4) for k, v in ms.items()},
                  "warp_instructions": {k: int(v) for k, v in inst.items()},
                  "_source": "%s: dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --metrics ... --clock-control none, "
                             "python bench.py --config %s --steps 2 --warmup 1); largest launch of each kernel" % (sys.argv[1], sys.argv[2])}, indent=1))

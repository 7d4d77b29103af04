"""profiles/r02_traffic.json from an ncu launch list (tools/gpu_final_r02.sh): DRAM bytes and time of the conv launches of the LAST pass.
usage: python tools/traffic_json.py gpurun_out/r02_launches_step.csv profiles/r02_traffic.json"""
import collections
import csv
import json
import sys

path, out = sys.argv[1], sys.argv[2]
lines = [l for l in open(path) if not l.startswith("==")]
by = collections.OrderedDict()
for row in csv.DictReader(lines):
    by.setdefault(int(row["ID"]), {"name": row["Kernel Name"]})[row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
ids = list(by)
# passes are identical: the last pass starts at the last pack_input / first kernel name repeat
first = by[ids[0]]["name"]
starts = [i for i in ids if by[i]["name"] == first]
last = [i for i in ids if i >= starts[-1]]
conv = [i for i in last if any(k in by[i]["name"] for k in ("conv_flat", "conv_pair", "res2_chain", "conv_umma"))]
byk = collections.Counter()
for i in conv:
    byk[by[i]["name"].split("<")[0].replace("void ", "").replace("svx::", "")] += 1
tot = lambda sel, m: sum(by[i].get(m, 0.0) for i in sel)
d = {
    "source": "%s (ncu launch list of tools/prof_step.py, last pass: 256 utterances x 200 frames x 80 bins, fp16)" % path.replace("gpurun_out", "profiles"),
    "conv_launches": len(conv),
    "conv_launches_by_kernel": dict(byk),
    "conv_dram_bytes_per_step": tot(conv, "dram__bytes_read.sum") + tot(conv, "dram__bytes_write.sum"),
    "all_dram_bytes_per_step": tot(last, "dram__bytes_read.sum") + tot(last, "dram__bytes_write.sum"),
    "conv_time_ms_under_ncu": tot(conv, "gpu__time_duration.sum") / 1e6,
    "step_time_ms_under_ncu": tot(last, "gpu__time_duration.sum") / 1e6,
}
d["conv_share_of_step"] = d["conv_time_ms_under_ncu"] / d["step_time_ms_under_ncu"]
json.dump(d, open(out, "w"), indent=1)
print(json.dumps(d, indent=1))

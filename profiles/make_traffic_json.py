"""profiles/r2_cfg<N>_traffic.json from an `ncu --set full --page raw --csv` dump of one step (forward SpMM, SDDMM, value
gather, transposed SpMM in launch order): DRAM bytes per launch = dram__bytes_read.sum + dram__bytes_write.sum.
bench.py attaches the dominant kernel's figure to its line as `roofline.traffic`.

    python profiles/make_traffic_json.py <config id> <raw.csv> <summary file cited in the JSON>
"""
import csv, json, os, sys

cfg, raw, summary = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
def col(name): return hdr.index(name)
def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
seen_spmm = 0
out = {}
for r in rows[2:]:
    name = r[col("Kernel Name")]
    rd = to_bytes(r[col("dram__bytes_read.sum")], units[col("dram__bytes_read.sum")])
    wr = to_bytes(r[col("dram__bytes_write.sum")], units[col("dram__bytes_write.sum")])
    t = float(r[col("gpu__time_duration.sum")].replace(",", ""))
    tu = units[col("gpu__time_duration.sum")]
    t_us = t * {"ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "nsecond": 1e-3, "msecond": 1e3}.get(tu, 1)
    if "gather_values" in name: tag = "spmm_gradB_gather"
    elif "sddmm" in name: tag = "sddmm"
    elif "spmm" in name:
        tag = "spmm_fwd" if seen_spmm == 0 else "spmm_gradB"
        seen_spmm += 1
    else: continue
    if tag in out: continue
    out[tag] = {"dram_bytes": int(rd + wr), "time_us": round(t_us, 3), "kernel": name}
doc = {"source": f"ncu --set full --clock-control none (one launch of each kernel of a config-{cfg} step); summary: {summary}; "
                 "dram__bytes_read.sum + dram__bytes_write.sum per launch", "config": cfg, "n_gpus": 1, "kernels": out}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), f"r2_cfg{cfg}_traffic.json")
json.dump(doc, open(path, "w"), indent=1)
print(path, {k: (v["dram_bytes"], v["time_us"]) for k, v in out.items()})

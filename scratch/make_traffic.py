"""gpurun_out/<name>_raw.csv (ncu --page raw --csv of ONE launch) -> traffic.json
{workload: dram bytes read+written per launch}"""
import csv, json, sys, os
out = {}
for wl, name in [a.split("=") for a in sys.argv[1:]]:
    f = f"gpurun_out/{name}_raw.csv"
    if not os.path.exists(f):
        continue
    rows = list(csv.reader(open(f)))
    hdr, units = rows[0], rows[1]
    r = rows[2]
    def val(metric):
        i = hdr.index(metric)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    out[wl] = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
json.dump(out, open("gpurun_out/traffic.json", "w"), indent=1)
print(out)

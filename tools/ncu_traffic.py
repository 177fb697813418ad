"""profiles/ncu_traffic.json: mean DRAM bytes (read + write) per launch of each kernel, from an `ncu --set full` report.
bench.py copies the GEMM figure into roofline.traffic (it cannot run ncu itself)."""
import collections, csv, json, subprocess, sys

def main(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    ni, ri, wi, ti = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows[2:]:
        name = r[ni].split("(")[0].replace("void ", "")
        b = float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]]
        a = agg[name]; a[0] += 1; a[1] += b; a[2] += float(r[ti])
    res = {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0], "ncu_us_per_launch": v[2] / v[0]} for k, v in agg.items()}
    json.dump({"source": rep, "note": "ncu --set full --clock-control none (cold L2 per replay): upper bound on in-situ traffic",
               "kernels": res}, open(out, "w"), indent=1)
    for k, v in res.items():
        print("%-60s %3d launches %8.1f MB/launch %7.1f us" % (k[:60], v["launches"], v["dram_bytes_per_launch"] / 1e6, v["ncu_us_per_launch"]))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

"""Per GEMM class: time-weighted sm__pipe_tensor_cycles_active (% of peak sustained active) from an `ncu --set full` report.
The epilogue template argument names the class (0 = QKV store16, 1 = c_fc QuickGELU, 2 = residual reduce-add: out_proj has
K = width and runs ~half as long as c_proj with K = 4 * width, 4 = patch embedding)."""
import collections, csv, subprocess, sys


def main(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    ni = hdr.index("Kernel Name")
    ti = hdr.index("gpu__time_duration.sum")
    pi = hdr.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    li = hdr.index("lts__throughput.avg.pct_of_peak_sustained_elapsed")
    di = hdr.index("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")
    res2 = [float(r[ti]) for r in rows[2:] if "pair_kernel<2" in r[ni]]
    split = (min(res2) + max(res2)) / 2 if res2 else 0.0
    agg = collections.OrderedDict()
    for r in rows[2:]:
        n, t = r[ni], float(r[ti])
        if "pair_kernel<0" in n: cls = "gemm_qkv      (pair_kernel<0>, store16)"
        elif "pair_kernel<1" in n: cls = "gemm_fc1      (pair_kernel<1>, QuickGELU16)"
        elif "pair_kernel<2" in n: cls = ("gemm_out      (pair_kernel<2>, K = width)" if t < split else "gemm_fc2      (pair_kernel<2>, K = 4 width)")
        elif "pair_kernel<4" in n: cls = "gemm_patch    (pair_kernel<4>)"
        elif "attention_tcgen05" in n: cls = "attention     (attention_tcgen05_kernel)"
        elif "layernorm" in n: cls = "layernorm"
        else: cls = n[:40]
        a = agg.setdefault(cls, [0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += t * float(r[pi]); a[3] += t * float(r[li]); a[4] += t * float(r[di])
    print("# %s: time-weighted per class (ncu replays are cold-cache and clock-uncontrolled)" % rep)
    print("%-50s %4s %10s %14s %12s %12s" % ("class", "n", "us/launch", "tensor pipe %", "LTS % peak", "DRAM % peak"))
    gt = gp = 0.0
    for k, a in agg.items():
        print("%-50s %4d %10.1f %14.1f %12.1f %12.1f" % (k, a[0], a[1] / a[0], a[2] / a[1], a[3] / a[1], a[4] / a[1]))
        if k.startswith("gemm_"):
            gt += a[1]; gp += a[2]
    if gt:
        print("%-50s %4s %10s %14.1f" % ("all GEMM launches (time-weighted)", "", "", gp / gt))


if __name__ == "__main__":
    main(sys.argv[1])

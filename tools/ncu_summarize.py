"""Turn ncu CSV exports into the per-kernel summaries kept under profiles/ (kernel names as used by ffb_prof / bench.py).

  launches : python tools/ncu_summarize.py launches  LOG.csv  > profiles/rNN_ncu_launches_n1.csv
             LOG.csv = `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file LOG.csv <cmd>`
  full     : python tools/ncu_summarize.py full      RAW.csv  > profiles/rNN_ncu_full_<what>_kernels.csv
             RAW.csv = `ncu -i REP.ncu-rep --page raw --csv` of a `ncu --set full --clock-control none` capture

Not part of the product."""
import csv
import re
import sys
from collections import OrderedDict

MODES = {0: "c2c_rows", 1: "c2c_cols", 2: "r2c_rows", 3: "c2r_rows", 4: "c2c_cols_tw", 5: "c2c_cols"}


def prof_name(kernel: str) -> str:
    """`void fft_pow2_kernel<double, (int)-1, 4, 512, 1, 16, 16, 4>(...)` -> `fft_c2c_cols_tw_f64_N64_fwd`"""
    m = re.search(r"fft_pow2_kernel<(float|double), (?:\(int\))?(-?1), (\d+), \d+, \d+, \d+((?:, \d+)+)>", kernel)
    if m:
        n = 1
        for r in m.group(4).split(",")[1:]:
            n *= int(r)
        return f"fft_{MODES[int(m.group(3))]}_{'f64' if m.group(1) == 'double' else 'f32'}_N{n}_{'fwd' if m.group(2) == '-1' else 'inv'}"
    m = re.search(r"fft_cols_stream_kernel<(float|double), (?:\(int\))?(-?1), (?:\(bool\))?(\d|true|false), \d+((?:, \d+)+)>", kernel)
    if m:
        n = 1
        for r in m.group(4).split(",")[1:]:
            n *= int(r)
        return f"fft_c2c_cols_stream_{'f64' if m.group(1) == 'double' else 'f32'}_N{n}_{'fwd' if m.group(2) == '-1' else 'inv'}"
    m = re.search(r"fs_pass_kernel<(float|double), (?:\(int\))?(-?1), (?:\(bool\))?(\d|true|false), (?:\(bool\))?(\d|true|false), (?:ffb::)?FsPlan<((?:\(int\))?\d+(?:, (?:\(int\))?\d+)+)>", kernel)
    if m:
        nums = [int(x) for x in re.findall(r"\d+", m.group(5))]
        n = 1
        for r in nums[1:]:
            n *= r
        is_a = m.group(3) in ("1", "true")
        return f"fft_fs_{'a' if is_a else 'b'}_{'f64' if m.group(1) == 'double' else 'f32'}_N{n}_{'fwd' if m.group(2) == '-1' else 'inv'}"
    m = re.search(r"fs_pass_multi_kernel<(float|double), (?:\(int\))?(-?1), (?:ffb::)?FsPlan<((?:\(int\))?\d+(?:, (?:\(int\))?\d+)+)>", kernel)
    if m:
        nums = [int(x) for x in re.findall(r"\d+", m.group(3))]
        n = 1
        for r in nums[1:]:
            n *= r
        return f"fft_fs_am_{'f64' if m.group(1) == 'double' else 'f32'}_N{n}_{'fwd' if m.group(2) == '-1' else 'inv'}"
    m = re.search(r"fft_l2four_kernel<(float|double), (?:\(int\))?(-?1), (?:ffb::)?FsPlan<([^>]*)>, (?:ffb::)?FsPlan<([^>]*)>", kernel)
    if m:
        def prod(t):
            v = [int(x) for x in re.findall(r"\d+", t)]
            n = 1
            for r in v[1:]:
                n *= r
            return n
        return f"fft_l2four_{'f64' if m.group(1) == 'double' else 'f32'}_N{prod(m.group(3)) * prod(m.group(4))}_{'fwd' if m.group(2) == '-1' else 'inv'}"
    m = re.search(r"void (\w+(?:<.*>)?)\(", kernel)
    return (m.group(1) if m else kernel).replace(", ", "; ")


def _rows(path):
    lines = [l for l in open(path, newline="") if l.startswith('"')]
    return list(csv.reader(lines))


def launches(path):
    rows = _rows(path)
    hdr = rows[0]
    ik, im, iu, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
    agg = OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"_(fwd|inv)$", "", prof_name(r[ik]))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) * scale[r[iu]]
    total = sum(a[1] for a in agg.values())
    steady = sum(a[1] for k, a in agg.items() if k.startswith(("fft_", "stage_", "calcN", "vort_", "mul2_")))
    out = ["kernel,launches,total_us,share,share_steady,avg_us"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        st = us / steady if k.startswith(("fft_", "stage_", "calcN", "vort_", "mul2_")) and steady else 0.0
        out.append(f"{k},{n},{us:.1f},{us / total:.4f},{st:.4f},{us / n:.1f}")
    return out


def full(path):
    rows = _rows(path)
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, key, default=float("nan")):
        if key not in col:
            return default
        try:
            return float(r[col[key]].replace(",", ""))
        except ValueError:
            return default

    def mb(r, key):
        u = units[col[key]] if key in col else "byte"
        return val(r, key) * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1e-6)

    def us(r):
        u = units[col["gpu__time_duration.sum"]]
        return val(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)

    out = ["kernel,grid,block,regs,duration_us,dram_read_MB,dram_write_MB,traffic_MB,dram_pct_peak,l1tex_pct,lts_pct,issue_active_pct,warps_active_pct,top_stalls"]
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        rd, wr = mb(r, "dram__bytes_read.sum"), mb(r, "dram__bytes_write.sum")
        stalls = sorted(((val(r, h, 0.0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stall_cols), reverse=True)
        top = " ".join(f"{n}={v:.2f}" for v, n in stalls[:3] if n not in ("selected",))
        grid = r[col["Grid Size"]].replace(",", " ") if "Grid Size" in col else ""
        block = r[col["Block Size"]].replace(",", " ") if "Block Size" in col else ""
        out.append(f"{prof_name(r[col['Kernel Name']])},{grid},{block},{val(r, 'launch__registers_per_thread'):.0f},{us(r):.1f},{rd:.1f},{wr:.1f},{rd + wr:.1f},"
                   f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f},{val(r, 'l1tex__throughput.avg.pct_of_peak_sustained_active'):.1f},"
                   f"{val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f},{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f},"
                   f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f},{top}")
    return out


if __name__ == "__main__":
    if len(sys.argv) != 3 or sys.argv[1] not in ("launches", "full"):
        raise SystemExit(__doc__)
    print("\n".join(launches(sys.argv[2]) if sys.argv[1] == "launches" else full(sys.argv[2])))

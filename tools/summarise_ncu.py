"""Turn ncu outputs into the text summaries kept under profiles/.

  python tools/summarise_ncu.py launches <launches.csv>          per-kernel count / time / share / DRAM bytes
  python tools/summarise_ncu.py full <report.ncu-rep>            per-kernel key metrics from a --set full capture
"""
import csv, io, re, subprocess, sys
from collections import defaultdict


def short(name):
    m = re.search(r"(k_[a-z0-9_]+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:40]


def launches(path):
    rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 14 and r[0].isdigit()]
    per = defaultdict(lambda: defaultdict(float))
    ids = defaultdict(set)
    for r in rows:
        k = short(r[4])
        v = float(r[14].replace(",", ""))
        unit = r[13]
        if r[12] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)
        elif unit in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        per[k][r[12]] += v
        ids[k].add(r[0])
    total = sum(p["gpu__time_duration.sum"] for p in per.values())
    print("kernel,launches,total_us,avg_us,share_of_gpu_time,avg_dram_read_MB,avg_dram_write_MB")
    for k, p in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = len(ids[k])
        t = p["gpu__time_duration.sum"]
        print(f"{k},{n},{t:.1f},{t / n:.2f},{t / total:.4f},{p['dram__bytes_read.sum'] / n / 1e6:.3f},{p['dram__bytes_write.sum'] / n / 1e6:.3f}")


WANT = [
    ("gpu__time_duration.sum", "time"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__bytes.sum.per_second", "dram_rate"),
    ("FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu_%"),
    ("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "alu_cyc_%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
    ("launch__shared_mem_per_block_static", "static_smem"),
    ("launch__occupancy_limit_shared_mem", "lim_smem"),
    ("launch__occupancy_limit_registers", "lim_regs"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf"),
    ("smsp__inst_executed.sum", "inst"),
]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    col = {h: i for i, h in enumerate(hdr)}
    seen = defaultdict(int)
    print("kernel,launch," + ",".join(f"{lab}[{units[col[m]]}]" if m in col else lab for m, lab in WANT))
    for r in rd[2:]:
        k = short(r[col["Kernel Name"]])
        seen[k] += 1
        vals = [r[col[m]] if m in col else "n/a" for m, _ in WANT]
        print(f"{k},{seen[k]}," + ",".join(v.replace(",", "") for v in vals))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

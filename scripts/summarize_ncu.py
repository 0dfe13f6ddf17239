"""Turn ncu outputs in gpurun_out/ into the small text summaries committed under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_bench_launches.txt
    python scripts/summarize_ncu.py kernel   gpurun_out/prof_contract.ncu-rep profiles/r01_contract_ncu.txt
"""
import collections
import csv
import io
import re
import subprocess
import sys
from pathlib import Path

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
]


def to_ms(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1.0)


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        name = re.sub(r"<unnamed>::", "", name)[:100]
        tot[name] += to_ms(r["Metric Value"], r["Metric Unit"])
        cnt[name] += 1
    T = sum(tot.values())
    out = [f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({src})",
           f"# {len(rows)} launches captured, total device time {T:.2f} ms (cold-cache, serialised: compare shares)",
           f"{'ms':>10} {'share':>7} {'count':>6}  kernel"]
    for k, v in tot.most_common(40):
        out.append(f"{v:10.3f} {100 * v / T:6.1f}% {cnt[k]:6d}  {k}")
    Path(dst).write_text("\n".join(out) + "\n")
    print("\n".join(out[:16]))


def kernel(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = [f"# ncu --set full --clock-control none --import-source on  ({src})"]
    names = [r[hdr.index("Kernel Name")] for r in data]
    out.append("kernels: " + " | ".join(re.sub(r"\(.*", "", n) for n in names))
    for i, h in enumerate(hdr):
        if h in KEYS or ("tensor" in h.lower() and "pct" in h):
            out.append(f"{h} [{units[i]}]: {[r[i] for r in data]}")
    Path(dst).write_text("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    if mode == "launches":
        launches(src, dst)
    else:
        kernel(src, dst)

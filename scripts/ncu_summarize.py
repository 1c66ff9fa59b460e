"""Turns the ncu outputs of scripts/gpu_round.sh into the tracked summaries under profiles/<round>/.

    python scripts/ncu_summarize.py gpurun_out/<tag> profiles/<round>

* launches.csv (gpu__time_duration.sum of every launch) -> launches_step.csv (the rows of ONE denoise step: the
  launches between two euler_step_kernel launches) + launch_summary.md (per-kernel launches / total us / share);
* gemm.ncu-rep, others.ncu-rep (--set full) -> ncu_full_summary.json (per launch: duration, tensor-pipe %, DRAM
  bytes, L2 hit rate, registers) and profiles/roofline_traffic.json (mean DRAM bytes per GEMM launch, read by bench.py).
"""
import csv
import io
import json
import re
import subprocess
import sys
from pathlib import Path

src, dst = Path(sys.argv[1]), Path(sys.argv[2])
dst.mkdir(parents=True, exist_ok=True)


def short(name: str) -> str:
    m = re.match(r"(?:void )?(?:ltxb::)?([A-Za-z0-9_]+(?:<[^>]*>)?)", name)
    return m.group(1) if m else name


def launch_list() -> None:
    p = src / "launches.csv"
    if not p.exists():
        return
    lines = [ln for ln in p.read_text().splitlines() if ln.startswith('"')]
    rows = list(csv.DictReader(io.StringIO("\n".join(lines))))
    rows = [r for r in rows if r["Metric Name"] == "gpu__time_duration.sum"]
    ends = [i for i, r in enumerate(rows) if short(r["Kernel Name"]).startswith("euler_step")]
    if len(ends) >= 2:  # second step = the timed one (warm-up first)
        step = rows[ends[0] + 1: ends[1] + 1]
    else:
        step = rows
    with open(dst / "launches_step.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["id", "kernel", "block", "grid", "ns"])
        for r in step:
            w.writerow([r["ID"], short(r["Kernel Name"]), r["Block Size"], r["Grid Size"], r["Metric Value"]])
    agg = {}
    for r in step:
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) / 1e3
    total = sum(a[1] for a in agg.values())
    out = [f"ncu --metrics gpu__time_duration.sum --clock-control none, one eager denoise step ({len(step)} launches, "
           f"{total / 1e3:.2f} ms serialised, cold cache: compare SHARES, not absolutes)", "",
           "| kernel | launches | total us | share |", "|---|---|---|---|"]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {n} | {us:.1f} | {100 * us / total:.1f}% |")
    (dst / "launch_summary.md").write_text("\n".join(out) + "\n")
    print("\n".join(out))


WANT = {
    "gpu__time_duration.sum": "time_us",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pct_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "tensor_inst_pct_active",
    "dram__bytes_read.sum": "dram_read_MB",
    "dram__bytes_write.sum": "dram_write_MB",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "launch__registers_per_thread": "regs",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
}
UNIT = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def full_sets() -> None:
    out = []
    for rep in ("gemm", "others"):
        p = src / f"{rep}.ncu-rep"
        if not p.exists():
            continue
        res = subprocess.run(["ncu", "-i", str(p), "--page", "raw", "--csv"], capture_output=True, text=True)
        rows = list(csv.reader(io.StringIO(res.stdout)))
        if len(rows) < 3:
            print(f"{p}: no rows ({res.stderr[:200]})")
            continue
        head, units = rows[0], rows[1]
        for r in rows[2:]:
            d = {"kernel": short(r[head.index("Kernel Name")]), "grid": r[head.index("Grid Size")]}
            for m, key in WANT.items():
                if m in head:
                    i = head.index(m)
                    try:
                        v = float(r[i].replace(",", ""))
                    except ValueError:
                        continue
                    if key in ("time_us", "dram_read_MB", "dram_write_MB"):
                        v *= UNIT.get(units[i], 1.0)
                    d[key] = round(v, 2)
            out.append(d)
    if not out:
        return
    (dst / "ncu_full_summary.json").write_text(json.dumps(out, indent=1) + "\n")
    g = [d for d in out if d["kernel"].startswith("gemm_bf16") and "dram_read_MB" in d]
    if g:
        traffic = sum(d["dram_read_MB"] + d["dram_write_MB"] for d in g) / len(g) * 1e6
        (dst.parent / "roofline_traffic.json").write_text(json.dumps({
            "ltxb_gemm_bf16": traffic, "unit": "bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean of "
            f"{len(g)} consecutive GEMM launches of one block, ncu --set full, {dst.name})"}, indent=1) + "\n")
    for d in out:
        print(d)


launch_list()
full_sets()

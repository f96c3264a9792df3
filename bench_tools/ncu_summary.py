"""Summarise the evidence pack produced on a GPU box by bench_tools/make_evidence.sh (no GPU needed here):

    python -m bench_tools.ncu_summary [gpurun_out/evidence] [profiles/r2/evidence]

reads prof.ncu-rep with `ncu -i ... --page raw --csv`, writes <dst>/ncu_summary.md (one row per profiled launch:
duration, launch geometry, registers, shared memory, tensor-pipe / warp activity, DRAM traffic, top warp-stall
reasons), and copies the kernel census (launches.csv), the compute-sanitizer logs and the report itself."""
import csv
import io
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
COLS = [
    ("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__grid_size", "CTAs"),
    ("launch__block_size", "thr"), ("launch__cluster_size", "cluster"), ("launch__registers_per_thread", "regs"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem KB"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM thr %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
]
STALL_PREFIX = "smsp__average_warps_issue_stalled_"     # ..._<reason>_per_issue_active.ratio


def main(argv) -> int:
    src = Path(argv[0]) if argv else ROOT / "gpurun_out" / "evidence"
    dst = Path(argv[1]) if len(argv) > 1 else ROOT / "profiles" / "r2" / "evidence"
    dst.mkdir(parents=True, exist_ok=True)
    rep = src / "prof.ncu-rep"
    raw = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith(STALL_PREFIX) and h.endswith(".ratio")]
    head = "(git HEAD at capture: " + ((src / "head.txt").read_text().strip() or "uncommitted working tree") + ")"
    out = ["# ncu --set full captures of smoke() (one-shot ps mode: nothing persistent resident)", "", head, "",
           "Cold-cache, serialised single launches under the profiler: compare shares and counters, not absolute times "
           "(CUDA-event timings of warm steps are in the bench JSONs). The first `fused_step_kernel` row is the 0-step "
           "warm-up launch.", "",
           "| " + " | ".join(n for _, n in COLS) + " | top stalls (warps stalled per issue) |",
           "|" + "---|" * (len(COLS) + 1)]
    for r in data:
        cells = []
        for key, _ in COLS:
            v = r[idx[key]] if key in idx else "n/a"
            if key == "Kernel Name":
                v = v.split("(")[0].replace("void ", "")[:44]
            elif key in idx and units[idx[key]] == "byte" and v.replace(".", "").isdigit():
                v = f"{float(v) / 1024:.0f} KB"
            else:
                try:
                    v = f"{float(v):.2f}"
                except ValueError:
                    pass
            cells.append(v)
        stalls = []
        for h in stall_cols:
            try:
                stalls.append((float(r[idx[h]]), h[len(STALL_PREFIX):-len("_per_issue_active.ratio")]))
            except ValueError:
                pass
        stalls.sort(reverse=True)
        cells.append(", ".join(f"{n} {v:.1f}" for v, n in stalls[:3]))
        out.append("| " + " | ".join(cells) + " |")
    (dst / "ncu_summary.md").write_text("\n".join(out) + "\n")
    for name in ("launches.csv", "sanitizer_memcheck.log", "sanitizer_racecheck.log", "sanitizer_synccheck.log",
                 "census_stdout.log", "head.txt", "prof.ncu-rep"):
        if (src / name).exists():
            shutil.copy(src / name, dst / name)
    print((dst / "ncu_summary.md").read_text())
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))

"""Summarise ncu outputs into profiles/: (1) launch list CSV -> per-kernel totals and step share; (2) .ncu-rep raw page ->
the roofline-relevant metrics of one kernel.  Usage:
    python scripts/ncu_summary.py launches gpurun_out/r01_launches.csv
    python scripts/ncu_summary.py kernel gpurun_out/r01_tc_v3.ncu-rep
"""
import csv, io, subprocess, sys
from collections import OrderedDict, defaultdict

def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ii = hdr.index("ID")
    seq = []
    for r in rows[1:]:
        try: seq.append((int(r[ii]), r[ki], float(r[vi].replace(",", "")) / 1e3))
        except Exception: pass
    # the steps are the launches from the first k_prologue on
    first = next(i for i, (_, n, _) in enumerate(seq) if "k_prologue" in n)
    step = seq[first:]
    tot = defaultdict(float); cnt = defaultdict(int)
    for _, n, us in step:
        n = n.split("(")[0].replace("void ", "")
        tot[n] += us; cnt[n] += 1
    total = sum(tot.values())
    print(f"# launches inside the ELBO steps (from first k_prologue on): {len(step)}; total {total:.1f} us (ncu: cold-cache, serialised)")
    print(f"{'kernel':60s} {'n':>4s} {'total us':>12s} {'avg us':>10s} {'share':>8s}")
    for n in sorted(tot, key=lambda k: -tot[k]):
        print(f"{n[:60]:60s} {cnt[n]:4d} {tot[n]:12.1f} {tot[n]/cnt[n]:10.1f} {100*tot[n]/total:7.2f}%")
    setup = seq[:first]
    print(f"# setup (synthetic data generation etc.): {len(setup)} launches, {sum(u for *_, u in setup)/1e3:.1f} ms")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "sm__inst_executed.sum.per_cycle_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]

def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"# {name[:100]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w); print(f"{w:75s} {r[i]:>18s} {units[i]}")

if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])

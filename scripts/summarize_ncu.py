"""Turn ncu artefacts (gpurun_out/*.ncu-rep, launch-list CSVs) into the small text summaries committed under
profiles/.  Usage: python scripts/summarize_ncu.py <round-tag>"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__inst_executed.sum",
        "sm__inst_executed.sum.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    units = rows[1]
    out = []
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        out.append(d)
    return out


for rep in sorted(f for f in os.listdir(os.path.join(ROOT, "gpurun_out")) if f.startswith(tag) and f.endswith(".ncu-rep")):
    lines = [f"# ncu --set full --clock-control none --import-source on, report {rep} (B200, sm_100a)"]
    for k in raw(os.path.join(ROOT, "gpurun_out", rep)):
        lines.append(f"\n## kernel {k.get('Kernel Name', ('?', ''))[0]}  grid {k.get('Grid Size', ('?',''))[0]} block {k.get('Block Size', ('?',''))[0]}")
        for w in WANT:
            if w in k:
                lines.append(f"{w:90s} {k[w][0]:>16s} {k[w][1]}")
    open(os.path.join(out_dir, rep.replace(".ncu-rep", "_ncu_full.txt")), "w").write("\n".join(lines) + "\n")
    print("wrote", rep)

for f in sorted(f for f in os.listdir(os.path.join(ROOT, "gpurun_out")) if f.startswith(tag + "_launches") and f.endswith(".csv")):
    rows = [r for r in csv.reader(open(os.path.join(ROOT, "gpurun_out", f))) if len(r) > 10 and r[0] != "ID"]
    n_all = len(rows)
    rows = [r for r in rows if "qgpu::" in r[4]]      # the data generator's torch kernels are not part of a step
    agg = collections.OrderedDict()
    for r in rows:
        name = r[4].split("(")[0].replace("void ", "")
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[-1]) / 1e6
    tot = sum(a[1] for a in agg.values())
    # kernels that run once per TABLE (registration / first use: narrowing, statistics, dictionary encoding), i.e. before
    # bench.py's timed region; everything else runs once per step
    INGEST = ("k_narrow", "k_minmax", "k_dict_insert", "k_dict_codes", "k_str_maxlen", "k_copy_bits", "k_rebase", "k_take_bits")
    step_tot = sum(a[1] for nme, a in agg.items() if not any(k in nme for k in INGEST))
    lines = [f"# ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv: launch list of `bench.py --query Q --steps 2 --warmup 1 --no-e2e --no-cpu`",
             f"# {n_all} launches captured, {len(rows)} of them the library's (qgpu::*); the rest is the synthetic-data generator (torch)",
             "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes", f"# total {tot:.3f} ms over {len(rows)} launches",
             f"# 'step share' = share among the per-step kernels ({step_tot:.3f} ms): the figure to hold against bench.py's kernel_share_of_step;",
             "# kernels marked (ingest) run once per table before the timed region (narrowing, statistics, dictionary encoding)",
             f"{'kernel':60s} {'launches':>8s} {'total_ms':>10s} {'avg_ms':>9s} {'share':>7s} {'step share':>11s}"]
    for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ingest = any(k in name for k in INGEST)
        lines.append(f"{name[:60]:60s} {n:8d} {ms:10.4f} {ms / n:9.4f} {ms / tot:7.1%} " + ("   (ingest)" if ingest else f"{ms / max(step_tot, 1e-9):11.1%}"))
    open(os.path.join(out_dir, f.replace(".csv", "_summary.txt")), "w").write("\n".join(lines) + "\n")
    print("wrote", f)

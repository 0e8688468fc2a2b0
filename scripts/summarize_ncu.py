"""Turn ncu output into the markdown summaries kept under profiles/.

  python scripts/summarize_ncu.py launches <launch-list.csv> <title> > profiles/<name>_summary.md
  python scripts/summarize_ncu.py full <report.ncu-rep> <title> > profiles/<name>_summary.md

`launches` reads the CSV of `ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`;
`full` reads a `--set full` report through `ncu -i ... --page raw --csv`.
"""
import collections
import csv
import io
import re
import subprocess
import sys

STEP_KERNELS = ["nodes_pack_kernel", "region_bounds_kernel", "brick_classify_kernel", "brick_update_kernel", "brick_update_smem_kernel", "proj_exact_kernel"]
FULL_KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
             "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
             "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
             "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
             "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
             "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
             "smsp__inst_executed_op_global_red.sum", "smsp__inst_executed_op_global_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum",
             "launch__occupancy_limit_registers", "launch__shared_mem_per_block_dynamic",
             "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("<unnamed>::", "").replace("dfb::", "")
    return re.sub(r"\(.*", "", name)


def launches(path, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    per = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
        per.setdefault(short(r["Kernel Name"]), []).append(us)
    print("# %s" % title)
    print("# command: ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv   (cold-cache, serialised: compare SHARES)\n")
    step = [(k, v) for k, v in per.items() if any(k.startswith(s) for s in STEP_KERNELS)]
    tot = sum(sum(v) / len(v) for _, v in step)
    print("## per-launch averages of the kernels of one step (share of the step)\n\n| kernel | us / launch | share of step |\n|---|---|---|")
    for k, v in sorted(step, key=lambda kv: STEP_KERNELS.index([s for s in STEP_KERNELS if kv[0].startswith(s)][0])):
        print("| %s | %.1f | %.1f %% |" % (k, sum(v) / len(v), 100 * sum(v) / len(v) / tot))
    print("\nstep total (sum of per-launch averages): %.1f us   (bench.py CUDA-event step time: see BENCH json `ms_per_step`)\n" % tot)
    alltot = sum(sum(v) for v in per.values())
    print("## all launches of the run (setup included: kNN table / brick / region builds happen once per graph revision)\n\n| kernel | launches | total us | share |\n|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print("| %s | %d | %.1f | %.1f %% |" % (k, len(v), sum(v), 100 * sum(v) / alltot))


def full(path, title):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ni = hdr.index("Kernel Name")
    print("# %s" % title)
    print("# numbers are per launch; no tensor-pipe activity anywhere (HBM-bound integer/fp32/fp64 work)\n")
    for r in rows[2:]:
        print("## %s" % r[ni])
        for k in FULL_KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("- %s = %s %s" % (k, r[i], units[i]))
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])

"""Summarise an .ncu-rep: per-kernel key metrics (raw page) and the top source lines by executed instructions."""
import csv, collections, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'launch__waves_per_multiprocessor', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']
for w in want:
    for i, c in enumerate(h):
        if c == w:
            print("%-78s %-8s %s" % (c, units[i], [r[i][:28] for r in data]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
OURS = ("kernels.cu", "kmap.cuh", "khash.h")
cur = None; per = collections.OrderedDict(); hdr = None; fpath = None
for r in rows:
    if r and r[0] == "File Path": fpath = r[1]
    if r and r[0] == "Function Name": cur = r[1].split('(')[0]; per.setdefault(cur, {}); hdr = None; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr is None or cur is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: ln = int(d["Line No"])
    except: continue
    base = (fpath or "").split("/")[-1]
    if "nimble_aligner_b200" not in (fpath or ""): ln = 0; base = "(toolkit headers)"
    a = per[cur].setdefault((base, ln), [0.0, 0.0, 0.0, ""])
    if base != "(toolkit headers)": a[3] = (r[1] if len(r) > 1 else "").strip()[:100]   # source text as embedded in the report (--import-source on)
    def f(x):
        try: return float(x)
        except Exception: return 0.0
    a[0] += f(d["Instructions Executed"]); a[1] += f(d["Thread Instructions Executed"]); a[2] += f(d["# Samples"])
for k, out in per.items():
    tot = sum(a[0] for a in out.values()) or 1; tots = sum(a[2] for a in out.values()) or 1
    print("\n== %s: inst=%d samples=%d" % (k, tot, tots))
    for (fn, ln), (ie, te, smp, txt) in sorted(out.items(), key=lambda x: -x[1][0])[:topn]:
        print("%-11s %4d inst%%=%5.1f thr/inst=%5.1f samp%%=%5.1f | %s" % (fn[:11], ln, 100 * ie / tot, te / ie if ie else 0, 100 * smp / tots, txt))

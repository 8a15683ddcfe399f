"""Summarise an .ncu-rep (raw + source pages) into text: python tools/ncu_summary.py rep [topN]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg', 'sm__cycles_active.avg', 'sm__cycles_elapsed.max',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__average_warp_latency_per_inst_issued.ratio', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum']
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for k in keys:
        if k in hdr:
            print(f"  {k:75s} {vals[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h:
            v = float(vals[i] or 0)
            if v > 0.05: print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:30s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
ts = sum(f(r, "# Samples") for r in data); ti = sum(f(r, "Instructions Executed") for r in data)
print(f"total samples {ts:.0f} warp-inst {ti:.0f}")
order = sorted(range(len(data)), key=lambda i: -f(data[i], "# Samples"))[:topn]
for i in sorted(order):
    r = data[i]
    print("%5d %5.1f%% inst=%11d thr/inst=%5.1f  %s" % (i, 100 * f(r, "# Samples") / ts, f(r, "Instructions Executed"), f(r, "Avg. Threads Executed"), r[ix["Source"]][:100]))
# coarse histogram of executed instructions by 100-instruction bucket
print("bucket  %inst  %samples  avg-threads")
B = 100
for b in range(0, len(data), B):
    seg = data[b:b + B]
    ii = sum(f(r, "Instructions Executed") for r in seg); ss = sum(f(r, "# Samples") for r in seg)
    tt = sum(f(r, "Thread Instructions Executed") for r in seg)
    if ii / ti > 0.01 or ss / ts > 0.01:
        print(f"{b:5d}  {100*ii/ti:5.1f}  {100*ss/ts:5.1f}  {tt/max(ii,1):5.1f}")

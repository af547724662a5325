"""Summarise an .ncu-rep: headline metrics, opcode mix and hottest SASS per kernel (run where ncu is)."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
nwarps = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__maximum_warps_per_active_cycle_pct']
stalls = [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')]
for r in rows[2:]:
    print("==", r[hdr.index('Kernel Name')][:60])
    for w in want:
        if w in hdr:
            print(f"   {w} = {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    st = sorted(((float(r[hdr.index(h)].replace(',', '') or 0), h.replace('smsp__pcsamp_warps_issue_stalled_', '')) for h in stalls), reverse=True)
    tot = sum(v for v, _ in st) or 1
    print("   stalls: " + ", ".join(f"{n}={100*v/tot:.0f}%" for v, n in st[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
kern = None; hdr = None; data = collections.OrderedDict()
for r in rows:
    if r and r[0] == "Kernel Name": kern = r[1][:50]; data.setdefault(kern, []); continue
    if r and r[0] == "Address": hdr = r; continue
    if kern and hdr and len(r) == len(hdr): data[kern].append(r)
for k, rs in data.items():
    ie = hdr.index("Instructions Executed"); te = hdr.index("Thread Instructions Executed"); ss = hdr.index("# Samples")
    tot = sum(int(r[ie]) for r in rs); tots = sum(int(r[ss]) for r in rs)
    print("==", k, "warp-instr", tot, "samples", tots, ("per-warp %.0f" % (tot / nwarps)) if nwarps else "")
    byop = collections.Counter(); bys = collections.Counter(); thr = collections.Counter()
    for r in rs:
        t = r[1].split()
        op = t[1] if t[0].startswith('@') else t[0]
        op = op.split('.')[0]
        byop[op] += int(r[ie]); bys[op] += int(r[ss]); thr[op] += int(r[te])
    for op, c in byop.most_common(18):
        print(f"  {op:8s} {100*c/tot:5.1f}% instr  {100*bys[op]/max(tots,1):5.1f}% samples  avgthr {thr[op]/max(c,1):5.1f}" + (f"  {c/nwarps:8.1f}/warp" if nwarps else ""))
    print("  hottest SASS by samples:")
    for r in sorted(rs, key=lambda r: -int(r[ss]))[:14]:
        print(f"    {int(r[ss]):6d}  exec {int(r[ie]):9d}  {r[1][:90]}")

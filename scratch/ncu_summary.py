import csv, collections, re, sys
raw, src, n = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows=list(csv.reader(open(raw))); hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__warps_eligible.avg.per_cycle_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
stalls=[h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
for r in rows[2:]:
    print('----', r[hdr.index('Kernel Name')][:70])
    for w in want:
        if w in hdr: i=hdr.index(w); print(f"  {w:75s} {r[i]} {units[i]}")
    st=sorted(((float(r[hdr.index(h)]),h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for h in stalls), reverse=True)[:7]
    print("  stalls/issue:", ", ".join(f"{b}={a:.2f}" for a,b in st))
rows=list(csv.reader(open(src)))
kern=None; hdr=None; agg={}
for r in rows:
    if len(r)>=2 and r[0]=="Kernel Name": kern=r[1]; agg[kern]=collections.Counter(); continue
    if r and r[0]=="Address": hdr=r; continue
    if kern and hdr and len(r)>5 and r[0].startswith("0x"):
        s=re.sub(r"^@!?U?P\d+\s+","",r[1].strip()); op=s.split()[0].split('.')[0]
        agg[kern][op]+=int(r[hdr.index("Instructions Executed")])
for k,c in agg.items():
    tot=sum(c.values()); cells=n*n
    print(k[12:80], "thread-inst/cell", round(tot*32/cells,1))
    print("   "+"  ".join(f"{op} {v*32/cells:.1f}" for op,v in c.most_common(26)))

import csv, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=None; data={}
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r)); key=(int(d['ID']), d['Kernel Name'][:26]); data.setdefault(key,{})[d['Metric Name']]=d['Metric Value']
tot={}
for k in sorted(data):
    m=data[k]
    t=float(m['gpu__time_duration.sum'])/1e6
    tot[k[1]]=tot.get(k[1],0)+t
    print(k[0], k[1], f"{t:7.2f} ms  inst {float(m['smsp__inst_executed.sum'])/1e9:6.2f}G issue {float(m['smsp__issue_active.avg.pct_of_peak_sustained_active']):5.1f}% lsb {float(m['smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio']):5.2f} thr {float(m['smsp__thread_inst_executed_per_inst_executed.ratio']):5.1f}")
print(tot, sum(tot.values()))

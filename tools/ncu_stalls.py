import csv, subprocess, sys
rep = sys.argv[1]; kid = sys.argv[2] if len(sys.argv) > 2 else "1"
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]; units=rows[1]; data=rows[2:]
idx={h:i for i,h in enumerate(hdr)}
for w in ['Grid Size','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','lts__t_sector_hit_rate.pct','launch__registers_per_thread','smsp__cycles_active.avg']:
    if w in idx: print('%-70s %-8s'%(w,units[idx[w]]), [d[idx[w]][:12] for d in data])
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-id","::regex:conv_umma:"+kid],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
hdr=rows[1]; data=rows[2:]; idx={h:i for i,h in enumerate(hdr)}
tot=sum(int(r[idx['# Samples']] or 0) for r in data)
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg={h:sum(int(r[idx[h]] or 0) for r in data) for h in stalls}
print('total samples',tot, sorted(agg.items(), key=lambda kv:-kv[1])[:6])
for r in sorted(data,key=lambda r:-int(r[idx['# Samples']] or 0))[:14]:
    s=int(r[idx['# Samples']]); best=max(stalls,key=lambda h:int(r[idx[h]] or 0))
    print('%6d %5.1f%% %-14s %s'%(s,100*s/tot,best,r[idx['Source']][:100]))

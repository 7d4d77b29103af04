import csv, collections, sys
path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches.csv'
lines=[l for l in open(path) if not l.startswith('==')]
r=list(csv.DictReader(lines))
by=collections.OrderedDict()
for row in r:
    by.setdefault(int(row['ID']),{'name':row['Kernel Name'][:22],'grid':row['Grid Size']})[row['Metric Name']]=float(row['Metric Value'].replace(',',''))
tot=0
lim=int(sys.argv[2]) if len(sys.argv)>2 else 30
for i,(k,v) in enumerate(by.items()):
    t=v.get('gpu__time_duration.sum',0)/1e3; rd=v.get('dram__bytes_read.sum',0); wr=v.get('dram__bytes_write.sum',0)
    tot+=t
    if i<lim: print(i,v['name'],v['grid'],'%.1f us'%t,'rd %.0fMB wr %.0fMB'%(rd/1e6,wr/1e6), 'GB/s %.0f'%((rd+wr)/ (t*1e-6)/1e9 if t else 0))
print('total ms',tot/1e3)

import csv,collections,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
tot=sum(int(r[ix['# Samples']]) for r in data)
ex=collections.Counter(); smp=collections.Counter(); stall=collections.Counter()
for r in data:
    n=int(r[ix['# Samples']]); e=int(r[ix['Instructions Executed']])
    src=r[ix['Source']].strip()
    op=(src.split()[0] if not src.startswith('@') else src.split()[1]).split('.')[0]
    ex[op]+=e; smp[op]+=n
    for k in hdr:
        if k.startswith('stall_') and '(Not' not in k and r[ix[k]].isdigit():
            stall[k]+=int(r[ix[k]])
print('total samples',tot,'executed',sum(ex.values()))
print('stalls',{k.replace('stall_',''):round(100*v/tot,1) for k,v in stall.most_common(9)})
print('sample ops',{k:round(100*v/tot,1) for k,v in smp.most_common(14)})
print('exec ops (M)',{k:round(v/1e6,1) for k,v in ex.most_common(24)})
top=sorted(data,key=lambda r:-int(r[ix['# Samples']]))[:int(sys.argv[2]) if len(sys.argv)>2 else 16]
for r in top:
    st={k:int(r[ix[k]]) for k in hdr if k.startswith('stall_') and '(Not' not in k and r[ix[k]].isdigit() and int(r[ix[k]])>0}
    print('%5.2f%% %-64s %s'%(100*int(r[ix['# Samples']])/tot, r[ix['Source']].strip()[:64], sorted(st.items(),key=lambda kv:-kv[1])[:2]))
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); h=rows[0]; d=rows[2]
for k,v in zip(h,d):
    if k in ('gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__warps_active.avg.per_cycle_active','sm__cycles_elapsed.max','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct'): print(k,v)

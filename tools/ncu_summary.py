import csv, re, collections, sys
raw, src = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0      # which captured launch
rows=list(csv.reader(open(raw)))
hdr=rows[0]; units=rows[1]; r=rows[2 + which]
print('kernel:', r[hdr.index('Kernel Name')])
want=['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','launch__block_size','smsp__inst_executed.sum','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','launch__shared_mem_per_block_dynamic','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active']
for i,h in enumerate(hdr):
    if h in want: print('  %-75s %s %s'%(h, r[i], units[i]))
rows=list(csv.reader(open(src)))
sections=[]; cur=None
for rr in rows:
    if rr and rr[0]=='Kernel Name': cur={'hdr':None,'rows':[]}; sections.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=rr; continue
    cur['rows'].append(rr)
nlaunch=max(1,len(list(csv.reader(open(raw))))-2)
sec=sections[which*max(1,len(sections)//nlaunch)]; hdr=sec['hdr']; ix={h:i for i,h in enumerate(hdr)}
def f(x):
    try: return float(x)
    except: return 0.0
op=collections.Counter(); samples=collections.Counter(); tot=0; stot=0
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
stalls=collections.Counter()
data=[rr for rr in sec['rows'] if len(rr)>=len(hdr)]
for rr in data:
    s_=rr[ix['Source']]; n=f(rr[ix['Instructions Executed']]); s=f(rr[ix['# Samples']])
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', s_)
    o=m.group(2).split('.')[0] if m else '?'
    op[o]+=n; samples[o]+=s; tot+=n; stot+=s
    for c in stall_cols: stalls[c]+=f(rr[ix[c]])
print('total inst %.3e samples %d'%(tot,stot))
for o,n in op.most_common(18): print('%-10s %12.3e %5.1f%%  samples %5.1f%%'%(o,n,100*n/tot,100*samples[o]/stot))
for c,v in stalls.most_common(8): print('%-25s %5.1f%%'%(c,100*v/stot))
idx=sorted(range(len(data)), key=lambda i:-f(data[i][ix['# Samples']]))[:14]
for i in sorted(idx):
    rr=data[i]; st={c:f(rr[ix[c]]) for c in stall_cols}; top=max(st,key=st.get)
    print('%5d %5.2f%% %-14s %s'%(i, 100*f(rr[ix['# Samples']])/stot, top, rr[ix['Source']][:80]))

#!/usr/bin/env python
"""Summarise an .ncu-rep here (no GPU needed): key metrics per kernel + top stall sites.
Usage: python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep [n_top]"""
import csv, subprocess, sys, io
rep=sys.argv[1]; ntop=int(sys.argv[2]) if len(sys.argv)>2 else 14
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr,units=rows[0],rows[1]
want=['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit','sm__warps_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum','launch__waves_per_multiprocessor','sm__cycles_elapsed.max']
idx=[i for i,h in enumerate(hdr) if h in want or h.startswith('launch__occupancy_limit')]
for r in rows[2:]:
    print('=' * 100)
    for i in idx:
        if i < len(r): print(f"  {hdr[i]:75s} {r[i]:>18s} {units[i]}")
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
kern=[];cur=None
for r in rows:
    if r and r[0]=="Kernel Name": cur={'name':r[1],'hdr':None,'rows':[]};kern.append(cur);continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=r;continue
    cur['rows'].append(r)
seen=set()
for k in kern:
    if k['name'] in seen: continue
    seen.add(k['name'])
    h=k['hdr'];si=h.index("# Samples");srci=h.index("Source");ie=h.index("Instructions Executed")
    sc=[i for i,x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
    tot=sum(int(r[si] or 0) for r in k['rows']) or 1
    print("\n##### stall sites:",k['name'][:90],"samples",tot,"sass instrs",len(k['rows']))
    agg={}
    for r in k['rows']:
        for i in sc:
            v=int(r[i] or 0)
            if v: agg[h[i]]=agg.get(h[i],0)+v
    print("   by reason:",{a:f"{100*b/tot:.1f}%" for a,b in sorted(agg.items(),key=lambda kv:-kv[1])[:8]})
    for r in sorted(k['rows'],key=lambda r:-int(r[si] or 0))[:ntop]:
        st={h[i]:int(r[i] or 0) for i in sc if int(r[i] or 0)>0}
        st=dict(sorted(st.items(),key=lambda kv:-kv[1])[:2])
        print(f"   {100*int(r[si])/tot:5.1f}% exec={r[ie]:>8} {r[srci].strip()[:64]:64s} {st}")

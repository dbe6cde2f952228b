import csv,collections,subprocess,sys
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr,units,vals=rows[0],rows[1],rows[2]
want=['gpu__time_duration.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__warps_eligible.avg.per_cycle_active','smsp__warps_active.avg.per_cycle_active','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct']
for h,u,v in zip(hdr,units,vals):
    if h in want: print(h,u,v)
st={h:float(v) for h,v in zip(hdr,vals) if 'smsp__average_warp' in h and 'issue_stalled' in h and h.endswith('_per_warp_active.pct') and v}
for h,v in sorted(st.items(), key=lambda kv:-kv[1])[:8]: print("  stall",h.replace('smsp__average_warps_issue_stalled_','').replace('_per_warp_active.pct',''),v)
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
agg=collections.OrderedDict(); cur=None; fname=None; tot=0; totsamp=0
for r in rows:
    if not r: continue
    if r[0]=="File Path": fname=r[1].split('/')[-1]; continue
    if r[0] in("Function Name","Line No"): continue
    if r[0]!="": cur=(fname,int(r[0]),r[1].strip()); agg.setdefault(cur,[0,0]); continue
    try: inst=int(r[7]); samp=int(r[4])
    except: continue
    agg[cur][0]+=inst; agg[cur][1]+=samp; tot+=inst; totsamp+=samp
print("total inst",tot,"samples",totsamp)
for (f,l,s),(inst,samp) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:topn]:
    print(f"{inst/tot*100:5.1f}% inst {samp/totsamp*100:5.1f}% samp  {f}:{l}  {s[:100]}")

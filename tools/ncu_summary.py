"""Helper: summarise an ncu report (per-kernel key metrics + opcode histogram)."""
import csv, collections, subprocess, sys, io
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw)))
hdr=rows[0]
def col(name): return hdr.index(name)
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__issue_active.avg.pct_of_peak_sustained_active','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__thread_inst_executed_per_inst_executed.ratio','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print(r[col('Kernel Name')][:60])
    for k in keys:
        try: print('   ',k, r[col(k)])
        except ValueError: print('   ',k,'n/a')
if len(sys.argv)>2:
    for kn in sys.argv[2:]:
        src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--kernel-name','regex:'+kn],capture_output=True,text=True).stdout
        rr=list(csv.reader(io.StringIO(src)))
        # may contain several kernels; take first block
        h=rr[1]; isrc=h.index('Source'); iex=h.index('Instructions Executed'); isamp=h.index('# Samples')
        seen=set(); ops=collections.Counter(); samp=collections.Counter(); tot=0
        for r in rr[2:]:
            if len(r)<=iex or r[0]=='Kernel Name' or r[0]=='Address': break
            key=(r[0],r[isrc])
            if key in seen: continue
            seen.add(key)
            try: n=int(r[iex]); s=int(r[isamp])
            except: continue
            t=r[isrc].split()
            if not t: continue
            op=(t[1] if t[0].startswith('@') else t[0]).split('.')[0]
            ops[op]+=n; samp[op]+=s; tot+=n
        print('==',kn,'total warp instr',tot)
        for op,n in ops.most_common(16): print(f'   {op:10s} {n:12d} {100*n/tot:5.1f}%  samples {samp[op]}')

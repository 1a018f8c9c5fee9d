"""Helper: per-kernel warp stall breakdown + memory summary from an ncu report."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def f(v):
    try: return float(v.replace(',', ''))
    except Exception: return float('nan')
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d['Kernel Name'][:70], ' grid', d.get('Grid Size'), 'block', d.get('Block Size'))
    st = [(h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), f(v))
          for h, v in d.items() if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
    st.sort(key=lambda x: -x[1])
    print('   stalls(warps per issue):', ', '.join('%s %.2f' % x for x in st[:6]))
    t = f(d['gpu__time_duration.sum'])
    rd, wr = f(d['dram__bytes_read.sum']), f(d['dram__bytes_write.sum'])
    print('   time %s %s  dram rd %s wr %s %s  dram%% %s' % (d['gpu__time_duration.sum'], '', d['dram__bytes_read.sum'], d['dram__bytes_write.sum'], '',
          d.get('dram__throughput.avg.pct_of_peak_sustained_elapsed', d.get('FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed'))))
    for k in ['sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
              'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
              'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
              'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
              'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
              'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps']:
        if k in d: print('   %-70s %s' % (k, d[k]))

import csv, sys, subprocess
rep = sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]; units = rows[1]; data = rows[2:]
want = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','smsp__inst_executed_op_shared_atom.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct']
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w:80s} {units[i]:10s} " + "  ".join(r[i] for r in data))
st=[]
for i,h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
        try: st.append((float(data[0][i]), h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')))
        except: pass
print('stalls:', ', '.join(f"{h}={v:.2f}" for v,h in sorted(st, reverse=True)[:9]))

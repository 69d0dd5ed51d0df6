#!/usr/bin/env python
"""Condensed text summary of an .ncu-rep (ncu --set full): duration, tensor-pipe / DRAM / L2 utilisation, traffic, stalls.
usage: ncu_summary.py file.ncu-rep"""
import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]
want=['gpu__time_duration.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','smsp__inst_executed.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max','smsp__cycles_active.avg','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warp_latency_per_inst_issued.ratio','sm__sass_thread_inst_executed_op_ffma_pred_on.sum']
for r in rows[2:]:
    d=dict(zip(hdr,r))
    print(d['Kernel Name'][:70], '| grid', d.get('launch__grid_size'), 'block', d.get('launch__block_size'))
    for k in want:
        if k in d: print('   %-75s %s %s'%(k,d[k],units[hdr.index(k)]))
    st=[(k,float(d[k])) for k in hdr if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('_per_issue_active.ratio') and d[k] not in('','no data')]
    st.sort(key=lambda x:-x[1])
    print('   stalls:', ', '.join('%s=%.2f'%(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v) for k,v in st[:6]))

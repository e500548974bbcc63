#!/usr/bin/env python
"""Tabulate a `ncu --metrics ... --csv --log-file` capture (tools/prof_round0.sh) per launch."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
byid = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = (int(row['ID']), row['Kernel Name'].split('(')[0][-22:])
    byid.setdefault(k, {})[row['Metric Name']] = float(row['Metric Value'].replace(',', ''))
short = {'dram__bytes_read.sum': 'dR_MB', 'dram__bytes_write.sum': 'dW_MB', 'gpu__time_duration.sum': 'us', 'smsp__inst_executed.sum': 'Minst',
         'smsp__thread_inst_executed_per_inst_executed.ratio': 'lanes', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'iss%',
         'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'lsb', 'lts__t_sectors_op_read.sum': 'l2R_MB',
         'lts__t_sectors_op_write.sum': 'l2W_MB', 'lts__t_sector_hit_rate.pct': 'l2hit', 'l1tex__t_sector_hit_rate.pct': 'l1hit',
         'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%', 'launch__registers_per_thread': 'regs',
         'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum': 'gld_MB', 'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum': 'lld_MB',
         'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum': 'lst_MB', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum': 'gst_MB'}
cols = ['us', 'Minst', 'lanes', 'iss%', 'lsb', 'occ%', 'regs', 'dR_MB', 'dW_MB', 'l2R_MB', 'l2W_MB', 'l2hit', 'l1hit', 'gld_MB', 'gst_MB', 'lld_MB', 'lst_MB']
print('%-22s' % 'kernel', ' '.join('%8s' % c for c in cols))
tot = 0
for k, v in byid.items():
    if 'reset' in k[1] or 'init' in k[1]:
        continue
    o = {}
    for m, s in short.items():
        x = v.get(m, float('nan'))
        if s in ('dR_MB', 'dW_MB'): x /= 1e6
        elif s.endswith('_MB'): x = x * 32 / 1e6
        elif s == 'us': x /= 1e3
        elif s == 'Minst': x /= 1e6
        o[s] = x
    tot += o['us']
    print('%-22s' % (str(k[0]) + ' ' + k[1][-14:]), ' '.join('%8.1f' % o[c] for c in cols))
print('total us', tot)

#!/usr/bin/env python
"""One-paragraph summary of an .ncu-rep (ncu --set full): duration, DRAM traffic, tensor-pipe and SM activity,
registers, the top stall reasons.  usage: ncu_summary.py report.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__cluster_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__cycles_elapsed.max', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_xu.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum']
for rep in sys.argv[1:]:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
    print(f'== {rep}: {kname[:60]}')
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            print(f'   {h:72s} {v} {u}')

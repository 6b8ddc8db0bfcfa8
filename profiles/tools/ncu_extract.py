#!/usr/bin/env python
"""Extract the metrics DESIGN.md / bench.py quote from one `ncu --set full` report into a small JSON.
usage: ncu_extract.py report.ncu-rep out.json [--stamp] [--rays N]
--stamp adds "source_sha16" (hash of ray_core.h + synthpy_b200.cu as they are NOW): bench.py attaches a capture to its
roofline only while that hash matches, so stale evidence cannot ride on a changed kernel.  Stamp a capture only when the
sources are the ones it was taken from."""
import csv, hashlib, io, json, os, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['Kernel Name', 'dram__bytes.sum.per_second', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'launch__block_size',
        'launch__grid_size', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum',
        'lts__t_sectors.sum.per_second', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__occupancy_limit_registers', 'sm__maximum_warps_per_active_cycle_pct',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'launch__shared_mem_per_block_dynamic',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum']
d = {}
for h, u, v in zip(hdr, units, vals):
    if h in keep or h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio'):
        d[h] = [v, u]
if '--stamp' in sys.argv:
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..')
    hh = hashlib.sha256()
    for f in ('ray_core.h', 'synthpy_b200.cu'):
        hh.update(open(os.path.join(root, 'synthpy_b200', 'csrc', f), 'rb').read())
    d['source_sha16'] = hh.hexdigest()[:16]
if '--rays' in sys.argv:
    d['rays_per_launch'] = int(float(sys.argv[sys.argv.index('--rays') + 1]))      # rays of the profiled launch (traffic scales with it)
json.dump(d, open(out, 'w'), indent=1, sort_keys=True)
for k in sorted(d):
    print(k, d[k])

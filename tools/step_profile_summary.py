#!/usr/bin/env python
"""From one `ncu --set full` capture of tools/prof_step.py: profiles/rNN_step_ncu.txt (counters of every kernel of the step,
per-source-line stall samples and the L1TEX data-pipe shares of the two tcgen05 kernels) and profiles/rNN_dram_traffic.json
(dram bytes per launch, read by bench.py for roofline.traffic).   usage: step_profile_summary.py file.ncu-rep rNN"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
ki, ri, wi = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
names = [('radix_pass', 'sort'), ('tile_ranges_fix', 'tile_ranges'), ('render_fwd_tc', 'render_fwd'), ('render_bwd_pix', 'render_bwd_pix'),
         ('render_bwd_chan_tc', 'render_bwd_chan'), ('preprocess_bwd', 'preprocess_bwd'), ('adam_multi', 'adam'),
         ('preprocess_kernel', 'preprocess'), ('emit_keys', 'emit_keys')]
acc = {}
for r in rows[2:]:
    for k, v in names:
        if k in r[ki]:
            acc[v] = acc.get(v, 0) + int(round((float(r[ri]) + float(r[wi])) * 1e6))
            break
d = {"_note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (bytes; 'sort' = the three radix passes together) from one "
              "ncu --set full capture of a whole cfgB step at the end of the round (tools/prof_step.py under ncu, summarised in "
              "profiles/%s_step_ncu.txt). ncu flushes the caches before every kernel, so these are cold-cache figures: in the "
              "pipeline the binning kernels and part of the backward hand-off records are L2 hits." % tag}
for k in ('preprocess', 'emit_keys', 'sort', 'tile_ranges', 'render_fwd', 'render_bwd_pix', 'render_bwd_chan', 'preprocess_bwd', 'adam'):
    d[k] = acc[k]
json.dump(d, open(os.path.join(ROOT, 'profiles', tag + '_dram_traffic.json'), 'w'), indent=1)
out = [subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_summary.py'), rep], capture_output=True, text=True).stdout]
PIPE = ('l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum', 'l1tex__data_pipe_tc_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.avg', 'sm__cycles_elapsed.avg',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__t_requests_srcunit_tex_op_red.sum')
for kern in ('render_fwd_tc', 'render_bwd_pix', 'render_bwd_chan_tc'):
    out.append('\n== %s: L1TEX data pipe (shared-memory loads/stores, global reductions and the tensor core\'s operand reads share it) ==\n' % kern)
    for r in rows[2:]:
        if kern in r[ki]:
            for k in PIPE:
                if k in h:
                    out.append('  %-92s %s\n' % (k, r[h.index(k)]))
            break
    out.append('\n== %s: warp-stall samples per CUDA source line (tools/ncu_lines.py) ==\n' % kern)
    out.append(subprocess.run([sys.executable, os.path.join(ROOT, 'tools', 'ncu_lines.py'), rep, kern, '16'], capture_output=True, text=True).stdout)
open(os.path.join(ROOT, 'profiles', tag + '_step_ncu.txt'), 'w').write(''.join(out))
print(json.dumps({k: v for k, v in d.items() if k != '_note'}))

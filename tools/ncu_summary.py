"""Summarise an .ncu-rep (run where ncu is installed):  python tools/ncu_summary.py rep.ncu-rep [out.json]"""
import csv, json, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'launch__block_size', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum']
STALL = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio'
STALLS = ['long_scoreboard', 'short_scoreboard', 'math_pipe_throttle', 'mio_throttle', 'wait', 'barrier',
          'branch_resolving', 'lg_throttle', 'not_selected', 'no_instruction', 'dispatch_stall', 'drain', 'imc_miss', 'tex_throttle']
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
idx = {h: i for i, h in enumerate(rows[0])}
units = rows[1]
# ncu picks one unit per column for the whole report: normalise time to ms and bytes to MB
SCALE = {'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0, 'second': 1e3, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0,
         'Gbyte': 1e3}
out = []
for r in rows[2:]:
    d = {'id': r[idx['ID']], 'kernel': r[idx['Kernel Name']].split('(')[0].replace('void ', '').replace('lfs::', '')}
    for w in WANT:
        if w in idx:
            try:
                d[w] = float(r[idx[w]].replace(',', ''))
                if w.startswith('gpu__time_duration') or w.startswith('dram__bytes'):
                    d[w] *= SCALE.get(units[idx[w]], 1.0)
            except ValueError:
                d[w] = r[idx[w]]
    d['stalls'] = {}
    for s in STALLS:
        k = STALL % s
        if k in idx:
            try: d['stalls'][s] = round(float(r[idx[k]]), 3)
            except ValueError: pass
    out.append(d)
if len(sys.argv) > 2:
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
for d in out:
    t = d.get('gpu__time_duration.sum', 0)
    rd, wr = d.get('dram__bytes_read.sum', 0), d.get('dram__bytes_write.sum', 0)
    print(f"{d['id']:>3} {d['kernel'][:34]:34s} {t:8.4f} ms  dram R {rd:8.1f} W {wr:8.1f} MB  {(rd + wr) / max(t, 1e-9) / 1e3:6.2f} TB/s"
          f"  regs {int(d.get('launch__registers_per_thread', 0)):3d} occ {d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0):5.1f}%"
          f" issue {d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0):5.1f}% L2hit {d.get('lts__t_sector_hit_rate.pct', 0):5.1f}"
          f" ld sec/req {d.get('l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 0) / max(d.get('l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 1), 1):4.1f}"
          f" st sec/req {d.get('l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 0) / max(d.get('l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 1), 1):4.1f}")
    print('      stalls', {k: v for k, v in sorted(d['stalls'].items(), key=lambda kv: -kv[1])[:5]})

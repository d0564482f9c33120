#!/usr/bin/env python
"""Refresh profiles/raster_ncu_summary.json (the numbers bench.py quotes for roofline.traffic) from an `ncu --set full` report of
k_raster_fast captured with the CURRENT chord_b200/csrc/raster.cu:

    python profiles/tools/update_raster_summary.py gpurun_out/raster_fast.ncu-rep <commit> <details-file-name>

bench.py only quotes the traffic when the sha256 recorded here equals that of the raster.cu it runs with."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rep, commit, details = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    h, u, v = rows[0], rows[1], rows[2]

    def m(k):
        x = float(v[h.index(k)].replace(',', ''))
        return x * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1, 'ms': 1, 'us': 1e-3, 'ns': 1e-6}.get(u[h.index(k)], 1)
    path = os.path.join(ROOT, 'profiles', 'raster_ncu_summary.json')
    s = json.load(open(path))
    s.setdefault('history', {})[f"round{s.get('round', '?')}_{s.get('commit', '?')}"] = {
        k: s[k] for k in ('duration_ms_under_ncu', 'dram_bytes_per_launch', 'traffic_over_algorithmic', 'warp_instructions_per_triangle') if k in s}
    rd, wr = m('dram__bytes_read.sum'), m('dram__bytes_write.sum')
    inst = m('smsp__inst_executed.sum')
    s.update({'commit': commit, 'raster_cu_sha256': hashlib.sha256(open(os.path.join(ROOT, 'chord_b200', 'csrc', 'raster.cu'), 'rb').read()).hexdigest(),
              'source': f'profiles/{details} (ncu --set full --clock-control none --import-source on, S5 781250 meshlets x 128 tri, 3840x2160)',
              'duration_ms_under_ncu': m('gpu__time_duration.sum'), 'dram_bytes_read': int(rd), 'dram_bytes_write': int(wr), 'dram_bytes_per_launch': int(rd + wr),
              'traffic_over_algorithmic': (rd + wr) / s['algorithmic_bytes_per_launch'], 'warp_instructions': int(inst),
              'warp_instructions_per_triangle': inst / (s['meshlets_per_launch'] * 128.0),
              'avg_active_threads_per_instruction': m('smsp__thread_inst_executed_per_inst_executed.ratio'),
              # the L2 atomic unit (every covered sample is one 64-bit RED.MAX)
              'l2_red_sectors': int(m('lts__t_sectors_srcunit_tex_op_red.sum')),
              'l2_red_pct_of_peak_sustained': m('lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed'),
              'l2_atomic_input_cycles_pct_of_peak': m('lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed'),
              'issue_slots_busy_pct': m('sm__inst_issued.avg.pct_of_peak_sustained_active') if 'sm__inst_issued.avg.pct_of_peak_sustained_active' in h else s.get('issue_slots_busy_pct')})
    json.dump(s, open(path, 'w'), indent=1)
    subprocess.run(f'ncu -i {rep} --page details > {os.path.join(ROOT, "profiles", details)}', shell=True)
    print({k: s[k] for k in ('commit', 'dram_bytes_per_launch', 'traffic_over_algorithmic', 'warp_instructions_per_triangle')})


if __name__ == '__main__':
    main()

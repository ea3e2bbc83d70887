#!/usr/bin/env python
"""Summarise ncu outputs: `launches <csv>` (per-kernel time shares) or `raw <ncu-rep>` (key metrics)."""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'lts__t_sectors_srcunit_tex.sum', 'smsp__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
    H = rows[hdr]
    ki, vi, ui = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        v = v / 1e6 if u == 'ns' else v / 1e3 if u in ('us', 'usecond') else v * 1e3 if u in ('s', 'second') else v
        agg[r[ki][:90]][0] += 1
        agg[r[ki][:90]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot:.3f} ms over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v[1]:10.3f} ms {v[0]:4d}x {100 * v[1] / tot:5.1f}%  {k}")


def raw(path):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H, U = rows[0], rows[1]
    for V in rows[2:]:
        print('==', V[H.index('Kernel Name')][:100])
        for i, h in enumerate(H):
            if h in WANT:
                print(f'  {h:84s} {V[i]:>22s} {U[i]}')


if __name__ == '__main__':
    {'launches': launches, 'raw': raw}[sys.argv[1]](sys.argv[2])

"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/ (tracked).

    python tools/ncu_summary.py launches gpurun_out/launches_r01c.csv profiles/r01_launches_cfg3_eval.md
    python tools/ncu_summary.py full gpurun_out/prof_syrk_r01c.ncu-rep profiles/r01_ncu_gemm_syrk.md
"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

FULL_KEYS = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
             'sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
             'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
             'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
             'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
             'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
             'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
             'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg', 'sm__cycles_active.avg',
             'smsp__cycles_active.avg', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
             'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
             'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warp_latency_issue_stalled_barrier.ratio',
             'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__sass_inst_executed_op_shared_ld.sum')


def launches(src, dst):
    rows = []
    with open(src) as f:
        text = f.read()
    start = text.index('"ID"')
    for r in csv.DictReader(io.StringIO(text[start:])):
        if r.get('Metric Name') == 'gpu__time_duration.sum':
            name = r['Kernel Name'].split('(')[0].replace('void ', '').replace('rc::', '')
            rows.append((name, float(r['Metric Value']) * (1e-3 if r['Metric Unit'] in ('ns', 'nsecond') else 1.0), r['Grid Size']))
    agg = defaultdict(lambda: [0, 0.0])
    for name, us, _ in rows:
        agg[name][0] += 1
        agg[name][1] += us
    total = sum(v[1] for v in agg.values())
    with open(dst, 'w') as out:
        out.write(f'# ncu launch list ({src}): {len(rows)} launches, {total / 1e3:.2f} ms summed device time\n\n')
        out.write('Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n')
        out.write('| kernel | launches | total ms | share | mean us |\n|---|---:|---:|---:|---:|\n')
        for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            out.write(f'| `{name}` | {cnt} | {us / 1e3:.3f} | {100 * us / total:.1f}% | {us / cnt:.1f} |\n')
    print(open(dst).read())


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2:]
    with open(dst, 'w') as out:
        out.write(f'# ncu --set full summary of {src}\n\n')
        for v in vals:
            rec = dict(zip(hdr, v))
            out.write(f"## {rec.get('Kernel Name')}  grid {rec.get('Grid Size')} block {rec.get('Block Size')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for i, h in enumerate(hdr):
                if h in FULL_KEYS:
                    out.write(f'| {h} | {v[i]} | {units[i]} |\n')
            out.write('\n')
    print(open(dst).read())


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])

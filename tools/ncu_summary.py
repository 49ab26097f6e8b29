"""Print the metrics we care about from an .ncu-rep (run here, no GPU needed)."""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__sectors_read.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum.per_second',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors_srcunit_tex_lookup_hit.sum',
        'lts__t_sectors_srcunit_tex_lookup_miss.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__data_pipe_lsu_wavefronts_mem_lg.sum', 'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'smsp__inst_executed_op_global_ld.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct',
        'sm__warps_active.avg.per_cycle_active']
STALL = 'smsp__average_warps_issue_stalled_'


def main(path, out=None):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        lines.append('----')
        stalls = []
        for i, h in enumerate(hdr):
            if h in KEEP:
                lines.append(f'{h} [{units[i]}] = {r[i]}')
            if h.startswith(STALL) and h.endswith('_per_issue_active.ratio'):
                try:
                    stalls.append((float(r[i]), h[len(STALL):-len('_per_issue_active.ratio')]))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        lines.append('warp stall reasons (warps per issue-active cycle): ' + ', '.join(f'{n}={v:.2f}' for v, n in stalls[:8]))
    text = '\n'.join(lines)
    print(text)
    if out:
        open(out, 'a').write(text + '\n')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)

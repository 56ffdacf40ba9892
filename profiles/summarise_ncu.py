"""ncu report -> the per-kernel summary CSV committed under profiles/.

    ncu -i gpurun_out/<name>.ncu-rep --page raw --csv > /tmp/raw.csv && python profiles/summarise_ncu.py /tmp/raw.csv profiles/<name>_full_summary.csv
"""
import csv
import sys

WANT = ['Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'sm__inst_executed.sum.per_cycle_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__sass_inst_executed_op_utcmma.sum',
        'smsp__sass_inst_executed_op_tmem_ldt.sum', 'smsp__sass_inst_executed_op_tmem_stt.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_global_red.sum',
        'lts__t_sectors_srcunit_tex_op_red.sum', 'lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max']


def short(n):
    return n.replace('unnamed>::', '').replace('void rf::', '').replace('rf::', '').replace('void ', '').split('(')[0][:48]


def main(src, dst):
    rows = list(csv.reader(open(src)))
    start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr, units, data = rows[start], rows[start + 1], rows[start + 2:]
    idx = {h: i for i, h in enumerate(hdr)}
    out = [['metric', 'unit'] + [short(r[idx['Kernel Name']]) for r in data]]
    for w in WANT:
        if w in idx:
            out.append([w, units[idx[w]]] + [r[idx[w]] for r in data])
    with open(dst, 'w', newline='') as f:
        csv.writer(f).writerows(out)
    print(f"{len(data)} launches -> {dst}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])

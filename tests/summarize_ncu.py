"""Summarises `ncu --set full` reports (ncu -i X.ncu-rep --page raw --csv) into a small JSON: per captured launch the duration,
SM / tensor-pipe / L2 / DRAM figures the rooflines of DESIGN.md quote.  Usage: python tests/summarize_ncu.py out.json a.ncu-rep ..."""
import csv
import json
import subprocess
import sys

KEYS = {
    'duration_us': 'gpu__time_duration.sum',
    'sm_cycles_elapsed': 'sm__cycles_elapsed.avg',
    'sm_cycles_active': 'sm__cycles_active.avg',
    'sm_clock_ghz': 'sm__cycles_elapsed.avg.per_second',
    'tensor_pipe_active_pct_of_active': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'tensor_pipe_active_pct_of_elapsed': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
    'sm_throughput_pct': 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'issue_active_pct': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'warps_active_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'dram_bytes_read': 'dram__bytes_read.sum',
    'dram_bytes_write': 'dram__bytes_write.sum',
    'dram_throughput_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'l2_throughput_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l2_to_sm_read_bytes': 'lts__t_bytes_equiv_l1sectormiss_pipe_lsu_mem_global_op_ld.sum',
    'l1tex_throughput_pct': 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
    'regs': 'launch__registers_per_thread',
    'dyn_smem_kb': 'launch__shared_mem_per_block_dynamic',
    'cluster': 'launch__cluster_dim_x',
}


def load(path):
    txt = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(kernel=r[hdr.index('Kernel Name')][:70], grid=r[hdr.index('Grid Size')], block=r[hdr.index('Block Size')])
        for k, m in KEYS.items():
            if m in hdr:
                d[k] = r[hdr.index(m)] + ((' ' + units[hdr.index(m)]) if units[hdr.index(m)] else '')
        for i, h in enumerate(hdr):
            if h.startswith('lts__t_sectors_srcunit_tex_op_read.sum') or h == 'lts__t_bytes.sum' or h == 'l1tex__m_xbar2l1tex_read_bytes.sum':
                d[h] = r[i] + ' ' + units[i]
        out.append(d)
    return out


if __name__ == '__main__':
    res = {p.split('/')[-1]: load(p) for p in sys.argv[2:]}
    json.dump(dict(source='ncu --set full --clock-control none (tests/ncu_capture_r02.sh); units as ncu --page raw reports them', captures=res),
              open(sys.argv[1], 'w'), indent=1)
    for p, caps in res.items():
        print('==', p)
        for c in caps:
            print('  %-48s grid %-14s %8s us  tensor %6s %% of active  sm %5s %%  dram r/w %s / %s  l2 %s %%' % (
                c['kernel'][:48], c['grid'], c.get('duration_us', '?').split()[0][:7], c.get('tensor_pipe_active_pct_of_active', '?').split()[0][:5],
                c.get('sm_throughput_pct', '?').split()[0][:5], c.get('dram_bytes_read', '?'), c.get('dram_bytes_write', '?'),
                c.get('l2_throughput_pct', '?').split()[0][:5]))

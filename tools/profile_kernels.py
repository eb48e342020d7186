#!/usr/bin/env python
"""Per-kernel evidence of one training step from an `ncu --set full` capture of the CURRENT binaries.

On the GPU box (one GPU, eager step so that ncu sees plain launches):

    ncu --set full --clock-control none --import-source on -o gpurun_out/r2_step \
        python tools/profile_step.py --steps 1
    ncu -i gpurun_out/r2_step.ncu-rep --page raw --csv --print-units base > gpurun_out/r2_step_raw.csv

Here:   python tools/profile_kernels.py gpurun_out/r2_step_raw.csv profiles/r2_kernels [--config c2]

writes profiles/r2_kernels.json (what bench.py reads `roofline.traffic` from) and a .md table: per
kernel of the LAST profiled step its C-ABI entry point, duration, DRAM bytes read / written,
achieved DRAM GB/s against the measured HBM peak, L2 bytes, tensor-pipe activity, issue-slot
utilisation and the algorithmic bytes of bench.stage_work() where that function knows the call.
ncu's per-launch times are cold-cache and serialised: shares, not absolutes, are comparable with
the CUDA-event timings bench.py takes live."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# kernel-name fragment -> C-ABI entry point (ordinals count launches of the same entry in a step)
ENTRY = [
    ('set_ctl_kernel', 'tn_set_ctl'),
    ('elastic_field_kernel', 'tn_elastic_field'),
    ('elastic_noise_kernel', 'tn_elastic_noise'),
    ('elastic_warp_kernel', 'tn_elastic_warp'),
    ('small_fprop_kernel', 'tn_convpool_fprop_train'),
    ('small_bwd_kernel', 'tn_convpool_bwd'),
    ('small_wgrad_kernel', 'tn_convpool_bwd'),
    ('softmax_head_kernel', 'tn_softmax_head_fwd_bwd'),
    ('head_dw_kernel', 'tn_softmax_head_bwd_weights'),
    ('colsum_kernel', 'tn_dense_bwd_weights(db)'),
    ('sgd_step_kernel', 'tn_sgd_momentum_maxnorm_update'),
    ('maxnorm_cols_kernel', 'tn_sgd_momentum_maxnorm_update(maxnorm)'),
    ('maxnorm_rows_kernel', 'tn_sgd_momentum_maxnorm_update(maxnorm)'),
    ('conv_tc_wgrad_kernel', 'tn_conv2d_tc_wgrad'),
    ('conv_tc_kernel', 'tn_conv2d_tc_fprop/dgrad'),
    ('softmax_nll_kernel', 'tn_softmax_nll_fwd_bwd'),
]
# the dense products of a step in launch order: forward of each hidden layer, then per layer
# (backward) weights before data
GEMM_ORDER_1HIDDEN = ['tn_dense_fwd', 'tn_dense_bwd_weights', 'tn_dense_bwd_data']

M = {
    'dur': 'gpu__time_duration.sum',
    'rd': 'dram__bytes_read.sum',
    'wr': 'dram__bytes_write.sum',
    'l2': 'lts__t_bytes.sum',
    'tensor': 'sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
    'tensor2': 'sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active',
    'issue': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
    'inst': 'smsp__inst_executed.sum',
    'occ': 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'regs': 'launch__registers_per_thread',
    'smem': 'launch__shared_mem_per_block_dynamic',
}


def fnum(v):
    try:
        return float(v.replace(',', ''))
    except Exception:
        return None


def main():
    raw, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr = rows[0]
    col = {}
    for i, h in enumerate(hdr):
        col.setdefault(h, i)
    # some metrics carry a section prefix ("TPC.TriageCompute.<name>")
    def find(name):
        if name in col:
            return col[name]
        for h, i in col.items():
            if h.endswith('.' + name):
                return i
        return None
    idx = {k: find(v) for k, v in M.items()}
    kn, gs, bs = col['Kernel Name'], col['Grid Size'], col['Block Size']
    launches = []
    for r in rows[2:]:
        if len(r) <= kn or not r[kn]:
            continue
        d = {k: (fnum(r[i]) if i is not None and i < len(r) else None) for k, i in idx.items()}
        d.update(name=r[kn], grid=r[gs], block=r[bs])
        launches.append(d)
    # last step = from the last set_ctl / first elastic_warp launch to the end
    starts = [i for i, l in enumerate(launches) if 'set_ctl_kernel' in l['name']]
    if not starts:
        starts = [i for i, l in enumerate(launches) if 'elastic_warp_kernel' in l['name']]
    step = launches[starts[-1]:] if starts else launches
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        hbm = float(peaks['hbm_gbs'])
    except Exception:
        hbm = 6650.0
    try:
        git = subprocess.run(['git', '-C', ROOT, 'rev-parse', '--short', 'HEAD'], capture_output=True,
                             text=True).stdout.strip()
    except Exception:
        git = '?'
    count = {}
    gemm_i = 0
    kernels = []
    total = sum((l['dur'] or 0) for l in step)
    for l in step:
        entry = None
        if 'gemm_tc' in l['name'] or 'sgemm_kernel' in l['name']:
            entry = GEMM_ORDER_1HIDDEN[gemm_i % 3]
            gemm_i += 1
        else:
            for frag, e in ENTRY:
                if frag in l['name']:
                    entry = e
                    break
        o = count.get(entry, 0)
        count[entry] = o + 1
        dur_us = (l['dur'] or 0) / 1e3              # base unit: ns
        dram = (l['rd'] or 0) + (l['wr'] or 0)
        kernels.append({
            'kernel': l['name'].replace('void ', '').split('(')[0], 'grid': l['grid'], 'block': l['block'],
            'entry': entry, 'ordinal': o, 'duration_us': round(dur_us, 2),
            'share': round((l['dur'] or 0) / total, 4) if total else None,
            'dram_read_bytes': l['rd'], 'dram_write_bytes': l['wr'], 'dram_bytes': dram,
            'dram_gbs': round(dram / (dur_us * 1e-6) / 1e9, 1) if dur_us else None,
            'dram_frac_of_peak': round(dram / (dur_us * 1e-6) / 1e9 / hbm, 4) if dur_us else None,
            'l2_bytes': l['l2'],
            'tensor_pipe_pct': l['tensor'] if l['tensor'] is not None else l['tensor2'],
            'issue_active_pct': l['issue'], 'warp_inst': l['inst'], 'warps_active_pct': l['occ'],
            'registers': l['regs'], 'dyn_smem': l['smem'],
        })
    doc = {'git': git, 'source': os.path.basename(raw), 'hbm_peak_gbs': hbm,
           'note': 'ncu --set full --clock-control none, eager step, cold cache, serialised launches',
           'step_us_sum': round(total / 1e3, 1), 'kernels': kernels}
    with open(out + '.json', 'w') as f:
        json.dump(doc, f, indent=1)
    with open(out + '.md', 'w') as f:
        f.write('# Kernels of one training step (ncu --set full, git {})\n\n'.format(git))
        f.write('Source: `{}`; HBM peak {} GB/s (MEASURED_PEAKS.json). {}.\n\n'.format(
            os.path.basename(raw), hbm, doc['note']))
        f.write('| # | kernel | entry point | grid | us | share | DRAM rd MB | DRAM wr MB | DRAM GB/s | % HBM peak '
                '| L2 MB | tensor pipe % | issue % | regs |\n|' + '---|' * 14 + '\n')
        for i, k in enumerate(kernels):
            mb = lambda v: '-' if v is None else '{:.2f}'.format(v / 1e6)
            pc = lambda v: '-' if v is None else '{:.1f}'.format(v)
            f.write('| {} | `{}` | {}#{} | {} | {} | {:.1%} | {} | {} | {} | {} | {} | {} | {} | {} |\n'.format(
                i, k['kernel'], k['entry'], k['ordinal'], k['grid'], k['duration_us'], k['share'] or 0,
                mb(k['dram_read_bytes']), mb(k['dram_write_bytes']), k['dram_gbs'],
                pc(100 * k['dram_frac_of_peak']) if k['dram_frac_of_peak'] is not None else '-',
                mb(k['l2_bytes']), pc(k['tensor_pipe_pct']), pc(k['issue_active_pct']),
                int(k['registers']) if k['registers'] else '-'))
        f.write('\n{} launches, {} us summed.\n'.format(len(kernels), doc['step_us_sum']))
    print('wrote', out + '.json', out + '.md', len(kernels), 'kernels')


if __name__ == '__main__':
    main()

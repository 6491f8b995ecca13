#!/usr/bin/env python
"""From one `ncu --set full` capture of a single forward (27 launches) write
   profiles/<tag>_ncu_full_summary.csv   one row per launch with the metrics the rooflines quote
   profiles/ncu_traffic.json             dram__bytes_read.sum + dram__bytes_write.sum per launch, mean per kernel family
                                         (bench.py copies these into `roofline.traffic`)
usage: tools/ncu_traffic.py gpurun_out/rNN_full.ncu-rep r01_runNN"""
import csv, io, json, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u, d = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(h)}
# the five tc_conv_kernel launches of a forward, in order: front, pool, mid-res x2, transposed
TC_ORDER = ['tc_conv_front_k3s2', 'tc_conv_pool_k2s2', 'tc_conv_k3_streamed', 'tc_conv_k3_streamed', 'tc_conv_up_convT']
NAMES = [('coarse_project_kernel', 'coarse_project_kernel'), ('relayout', 'relayout_kernel'), ('gather_stream_kernel', 'gather_stream_kernel'),
         ('tc_norm_act_kernel', 'tc_norm_act_kernel'), ('tc_conv3', 'tc_conv3_stacked'), ('tc_head_centroid_kernel', 'tc_head_centroid_kernel'),
         ('centroid_finalize_kernel', 'centroid_finalize_kernel')]
KEEP = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active']


def num(r, k):
    v = float(r[ix[k]].replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u[ix[k]], 1)


acc, tc_seen = {}, 0
with open(f'profiles/{tag}_ncu_full_summary.csv', 'w') as f:
    w = csv.writer(f)
    w.writerow(['id', 'family', 'kernel', 'grid', 'block'] + KEEP)
    for i, r in enumerate(d):
        name = r[ix['Kernel Name']]
        fam = None
        if 'tc_conv_kernel' in name:
            fam = TC_ORDER[tc_seen % len(TC_ORDER)]; tc_seen += 1
        else:
            for key, fm in NAMES:
                if key in name:
                    fam = fm
                    break
        fam = fam or name
        w.writerow([i, fam, name[:60], r[ix['Grid Size']], r[ix['Block Size']]] + [r[ix[k]] + ' ' + u[ix[k]] for k in KEEP])
        acc.setdefault(fam, []).append(num(r, 'dram__bytes_read.sum') + num(r, 'dram__bytes_write.sum'))
tr = {k: int(sum(v) / len(v)) for k, v in acc.items()}
json.dump({"source": f"profiles/{tag}_ncu_full_summary.csv (ncu --set full --clock-control none, one forward of workload c3_full3d_example, "
                     f"B=32, {len(d)} launches): dram__bytes_read.sum + dram__bytes_write.sum, mean per launch of the kernel family",
           "dram_bytes_per_launch": tr}, open('profiles/ncu_traffic.json', 'w'), indent=1)
print(json.dumps(tr, indent=1))

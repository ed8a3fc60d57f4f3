"""Per-kernel-family DRAM traffic and device time per step from an ncu launch list taken with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv ...
over a window of whole training steps of bench.py (cold-cache, serialised launches: compare SHARES of time, the byte
counts are per launch). Usage: ncu_traffic.py launches.csv STEPS_IN_WINDOW|gemm:LAUNCHES_PER_STEP [out.json]"""
import collections
import csv
import json
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
steps_arg = sys.argv[2]   # a number, or gemm:<launches per step> to derive it from the window's GEMM launch count
h = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[h]
ki, mi, ui, vi, idi = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Unit'), hdr.index('Metric Value'), hdr.index('ID')
scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1.0, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6, 'nsecond': 1.0}
launch = collections.defaultdict(dict)
names = {}
for r in rows[h + 1:]:
    if len(r) <= vi:
        continue
    names[r[idi]] = re.sub(r'\(CUtensorMap.*|\(.*', '', r[ki]).replace('void ', '')
    launch[r[idi]][r[mi]] = float(r[vi].replace(',', '')) * scale.get(r[ui], 1.0)


def family(name: str) -> str:
    if 'gemm_kernel' in name or 'splitk' in name:
        return 'gemm'
    if 'gemm_ln' in name:
        return 'linear_ln_fused'
    if 'attn' in name:
        return 'attention'
    if 'ln_' in name or 'dropout_add' in name or 'reduce_partials' in name:
        return 'layernorm'
    if 'fbank' in name or 'cmvn' in name:
        return 'fbank'
    if 'adam' in name or 'sqnorm' in name or 'grad32' in name:
        return 'optimizer'
    if 'nccl' in name.lower():
        return 'nccl'
    return 'other'


agg = collections.defaultdict(lambda: {'launches': 0, 'ns': 0.0, 'read': 0.0, 'write': 0.0})
for lid, m in launch.items():
    a = agg[family(names[lid])]
    a['launches'] += 1
    a['ns'] += m.get('gpu__time_duration.sum', 0.0)
    a['read'] += m.get('dram__bytes_read.sum', 0.0)
    a['write'] += m.get('dram__bytes_write.sum', 0.0)
steps = agg['gemm']['launches'] / float(steps_arg[5:]) if steps_arg.startswith('gemm:') else float(steps_arg)
tot_ns = sum(a['ns'] for a in agg.values())
out = {'steps_in_window': steps, 'families': {}}
print(f'{"family":12s} {"launches/step":>13s} {"ms/step":>9s} {"share":>6s} {"DRAM read MB/step":>18s} {"DRAM write MB/step":>19s}')
for k, a in sorted(agg.items(), key=lambda x: -x[1]['ns']):
    row = {'launches_per_step': a['launches'] / steps, 'ms_per_step': a['ns'] / steps / 1e6, 'share_of_time': a['ns'] / tot_ns,
           'dram_read_bytes_per_step': a['read'] / steps, 'dram_write_bytes_per_step': a['write'] / steps}
    out['families'][k] = row
    print(f'{k:12s} {row["launches_per_step"]:13.1f} {row["ms_per_step"]:9.3f} {100 * row["share_of_time"]:5.1f}% '
          f'{row["dram_read_bytes_per_step"] / 1e6:18.1f} {row["dram_write_bytes_per_step"] / 1e6:19.1f}')
g = out['families'].get('gemm')
if g:
    out['gemm_dram_bytes_per_step'] = g['dram_read_bytes_per_step'] + g['dram_write_bytes_per_step']
    out['gemm_ms_per_step_ncu'] = g['ms_per_step']
f = out['families'].get('fbank')
if f:
    out['fbank_dram_bytes_per_step'] = f['dram_read_bytes_per_step'] + f['dram_write_bytes_per_step']
if len(sys.argv) > 3:
    json.dump(out, open(sys.argv[3], 'w'), indent=1)

"""Summarise an `ncu --set full` capture (exported with `ncu -i x.ncu-rep --page raw --csv`) per kernel class.

    python tools/ncu_summary.py raw.csv [traffic.json backbone]

Prints, per kernel class, launches, mean duration, DRAM read/write MB per launch, tensor-pipe active % (of active and of
elapsed cycles), issue-slot % and L2->SM read MB; with a second argument also merges the per-launch DRAM traffic
(dram__bytes_read.sum + dram__bytes_write.sum, mean over the class's launches) into the JSON file `bench.py` reads for
`roofline.traffic`.
"""
import collections
import csv
import json
import re
import sys

csv.field_size_limit(1 << 30)
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h, units = rows[hi], rows[hi + 1]
col = {n: i for i, n in enumerate(h)}

SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'usecond': 1.0,
         'nsecond': 1e-3, 'msecond': 1e3, 'second': 1e6}


def get(r, name, scaled=False):
    i = col.get(name)
    if i is None or i >= len(r) or r[i] in ('', 'n/a'):
        return float('nan')
    v = float(r[i].replace(',', ''))
    return v * SCALE.get(units[i], 1.0) if scaled else v


# ncu exports names like "void tc_conv_kernel<2>(Params)"; keep the template argument, drop the rest
def short(name):
    name = name.replace('void ', '')
    name = re.sub(r'\(.*$', '', name)
    name = name.replace('__nv_bfloat16', 'bf16')
    return name[:44]


# bench.py's names for the C entry points (the keys of profiles/traffic.json)
ENTRY = [('gbn_bwd', 'dards_gbn_bwd'), ('gbn_fwd', 'dards_gbn_fwd'), ('gbn_apply', 'dards_gbn_apply_fwd'),
         ('tc_conv_bn', 'dards_conv1d_bn_fwd'), ('tc_wgrad_kernel', 'dards_conv1d_wgrad_accum'),
         ('unpack_wgrad', 'dards_unpack_wgrad_batched'),
         ('tc_wgrad_reduce', 'wgrad_reduce'), ('tc_conv', 'dards_conv1d:tcgen05'), ('stem_fwd', 'dards_stem_fwd'), ('stem_bwd', 'dards_stem_bwd'),
         ('dropout', 'dards_dropout')]

agg = collections.OrderedDict()
fam = collections.OrderedDict()   # the same, template arguments dropped (one row per kernel family)
for r in rows[hi + 2:]:
    if len(r) <= col['Kernel Name']:
        continue
    k = short(r[col['Kernel Name']])
    a = agg.setdefault(k, collections.defaultdict(list))
    fa = fam.setdefault(re.sub(r'<.*$', '', k), collections.defaultdict(list))
    for dst in (fa,):
        dst['t'].append(get(r, 'gpu__time_duration.sum', True))
        dst['rd'].append(get(r, 'dram__bytes_read.sum', True))
        dst['wr'].append(get(r, 'dram__bytes_write.sum', True))
        dst['ta'].append(get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'))
        dst['te'].append(get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'))
        dst['is'].append(get(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'))
        dst['l2'].append(get(r, 'l1tex__m_xbar2l1tex_read_bytes.sum', True))
        dst['grid'].append(r[col['Grid Size']])
    a['t'].append(get(r, 'gpu__time_duration.sum', True))
    a['rd'].append(get(r, 'dram__bytes_read.sum', True))
    a['wr'].append(get(r, 'dram__bytes_write.sum', True))
    a['ta'].append(get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'))
    a['te'].append(get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'))
    a['is'].append(get(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'))
    a['l2'].append(get(r, 'l1tex__m_xbar2l1tex_read_bytes.sum', True))
    a['grid'].append(r[col['Grid Size']])


def mean(v):
    v = [x for x in v if x == x]
    return sum(v) / len(v) if v else float('nan')


HDR = "%-44s %4s %9s %9s %9s %8s %8s %7s %9s  %s" % ("kernel", "n", "time_us", "dram_rdMB", "dram_wrMB", "tens%act",
                                                      "tens%ela", "issue%", "l2->smMB", "grids")


def table(d):
    print(HDR)
    for k, a in sorted(d.items(), key=lambda kv: -sum(kv[1]['t'])):
        grids = sorted(set(a['grid']))
        print("%-44s %4d %9.2f %9.2f %9.2f %8.1f %8.1f %7.1f %9.1f  %s" % (
            k, len(a['t']), mean(a['t']), mean(a['rd']) / 1e6, mean(a['wr']) / 1e6, mean(a['ta']), mean(a['te']),
            mean(a['is']), mean(a['l2']) / 1e6, ' '.join(grids[:4])))


print("per kernel family (means per launch; tens%act = tensor pipe active, % of the cycles the SM was active):")
table(fam)
print()
print("per template instance:")
table(agg)
traffic = {}
for k, a in agg.items():
    for pat, entry in ENTRY:
        if pat in k:
            t = traffic.setdefault(entry, [])
            t.extend(x + y for x, y in zip(a['rd'], a['wr']))
            break

if len(sys.argv) > 3:
    path, backbone = sys.argv[2], sys.argv[3]
    try:
        doc = json.load(open(path))
    except FileNotFoundError:
        doc = {}
    d = doc.setdefault(backbone, {})
    for entry, v in traffic.items():
        if entry == 'wgrad_reduce':
            continue
        if entry == 'dards_conv1d_wgrad:tcgen05':  # one entry-point call = the split-K kernel + its reduce kernel
            d[entry] = mean(v) + (mean(traffic['wgrad_reduce']) if 'wgrad_reduce' in traffic else 0.0)
        elif entry == 'dards_conv1d:tcgen05':      # forward and dgrad share the kernel; the capture holds both
            d['dards_conv1d_fwd:tcgen05'] = d['dards_conv1d_dgrad:tcgen05'] = mean(v)
        else:
            d[entry] = mean(v)
    doc['_source'] = ("ncu --set full --clock-control none of `bench.py --steps 2 --warmup 3 --no-cpu --no-graph` "
                      "(tools/gpu_profile.sh), dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches of "
                      "each kernel class in one step; summary tables under profiles/")
    json.dump(doc, open(path, 'w'), indent=1)
    print("wrote", path)

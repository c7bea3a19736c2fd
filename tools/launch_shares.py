"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel family (last `frac` of the launches)."""
import csv, collections, sys
path = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
rows = list(csv.reader(open(path)))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr, data = rows[hi], rows[hi + 1:]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
part = data[int(len(data) * (1 - frac)):]
KEYS = ('conv_tc_kernel', 'wgrad_tc_kernel', 'wgrad_reduce', 'bn_reduce', 'bn_apply', 'bn_bwd_apply', 'bn_finalize', 'pwf_fwd_pass',
        'pwf_bwd_pass', 'pwf_bwd_finish', 'fusion_combine_bwd', 'fusion_kernel', 'grad_pack', 'pack_weights', 'nchw_to_nhwc', 'channel_sum',
        'act_unpack', 'bev_pack', 'add_f32', 'pwf_running', 'sums_to_f32', 'voxel', 'bev_scatter')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in part:
    name = r[kn]
    for key in KEYS:
        if key in name:
            name = key
            break
    else:
        name = 'torch:' + name[:48]
    v = float(r[mv].replace(',', ''))
    v = v / 1e3 if r[mu] == 'ns' else v
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{len(part)} launches, {tot:.1f} us in total (ncu per-launch durations: cold caches, serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{k:50s} n={v[0]:4d} {v[1]:10.1f} us {100 * v[1] / tot:5.1f}%")

"""Prints the main figures of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
for k in ('n_gpus', 'value', 'ms_per_step', 'stage_ms_per_step', 'clocks', 'gpu_launches'):
    print(k, json.dumps(d[k])[:300])
print('parity', {k: v for k, v in d['parity'].items() if k != 'how'})
e = d['e2e']
print('e2e', e['value'], e['ms_per_step'], e.get('call'), e['stage_ms'], e.get('host_threads'), e['sampled_pairs_exact'], e['matches_device_run'])
print('roofline', d['roofline']['kernel'][:60], {k: d['roofline'][k] for k in ('achieved', 'peak', 'frac', 'avg_launch_ms')})
r8 = d['roofline_int8']
print('int8', r8['kernel'][:60], {k: r8[k] for k in ('achieved', 'peak', 'frac', 'ms_per_step', 'frac_of_nominal_int8_peak')})
for k, v in d.get('extras', {}).items():
    if 'ms_per_step' in v:
        print(k, round(v['ms_per_step'], 3), '%.3e' % v['value'])
    else:
        print(k, json.dumps(v)[:600])

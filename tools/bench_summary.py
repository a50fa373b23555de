#!/usr/bin/env python
"""Pretty-print the JSON line(s) of bench.py from stdin or a file."""
import json, sys
src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for line in src:
    line = line.strip()
    if not line.startswith('{'):
        if line: print(line[:220])
        continue
    d = json.loads(line)
    if d.get('impl') == 'reference':
        print('REFERENCE', d.get('value'), d.get('unit'), d.get('cpu_baseline'))
        continue
    r = d['roofline']
    if d.get('e2e') is None:
        print(f"SHARDED value {d['value']:.0f} {d['unit']} n_gpus {d['n_gpus']} ms/step {d['ms_per_step']:.2f} roof {r['frac']:.3f} counts {d['config'].get('owned_suffixes_per_rank')} rounds {d['config'].get('rounds')} nccl_rx {d['config'].get('nccl_bytes_received_per_rank_per_step')}")
        for k, v in d['phases_rank0'].items(): print(f"   {k:12s} {v['ms_per_step']:8.3f} ms {v['launches_per_step']:5.1f}")
        continue
    print(f"value {d['value']:.0f} {d['unit']}  ms/step {d['ms_per_step']:.2f}  e2e {d['e2e']['value']:.0f} ({d['e2e']['ms_per_step']:.1f} ms)  "
          f"roof {r['frac']:.3f} ({r['achieved']:.0f} GB/s, {r['avg_launch_ms']:.3f} ms/launch, share {r['share_of_step']:.2f})  launches {d['gpu_launches']}")
    if d.get('cpu_baseline'): print('  cpu', round(d['cpu_baseline']['value'], 1), 'MB/s', d['cpu_baseline']['cores'], 'threads')
    print('  clocks', d.get('clocks'))
    for k, v in d['phases'].items():
        print(f"   {k:12s} {v['ms_per_step']:8.3f} ms  {v['launches_per_step']:5.1f} launches  {v['alg_GB_per_step']:7.2f} GB")

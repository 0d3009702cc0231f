"""Pretty-print a bench.py --dump-kernels file: python tools/kdump.py gpurun_out/kernelsN.json [bench_N.json] [rows]"""
import json
import sys

k = json.load(open(sys.argv[1]))
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 45
for r in k[:rows]:
    print(f"{r['name']:48s} {r['launches_per_step']:5.0f} {r['ms_per_step']:.3f} {1e3 * r['ms_per_step'] / r['launches_per_step']:7.1f}us "
          f"{'' if r['tflops'] is None else round(r['tflops'])}")
print("sum of kernels per step (eager, event-timed):", round(sum(r['ms_per_step'] for r in k), 3), "ms")
groups = {}
for r in k:
    key = r['name'].split('_p')[0] if r['name'].startswith('bn_') else r['name'].split('_')[0]
    groups[key] = groups.get(key, 0) + r['ms_per_step']
print({a: round(b, 3) for a, b in sorted(groups.items(), key=lambda kv: -kv[1])})
if len(sys.argv) > 2 and sys.argv[2] != '-':
    b = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print({x: b[x] for x in ['value', 'ms_per_step', 'e2e', 'roofline', 'gpu_launches']})

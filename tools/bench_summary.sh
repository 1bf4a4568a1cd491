#!/bin/bash
# usage: tools/bench_summary.sh <workload> [steps] [extra bench flags]  -- compact stage table (tools, not product)
python bench.py --workload $1 --steps ${2:-5} --warmup 3 --no-cpu-baseline ${@:3} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
e=d['e2e']['value'] if d.get('e2e') else float('nan')
print('$1', 'Mp/s=%.0f ms/step=%.3f e2e=%.0f' % (d['value'], d['ms_per_step'], e), {k: round(v,3) for k,v in d['stages']['ms'].items()}, 'roof', d['roofline']['kernel'], round(d['roofline']['frac'],3), 'bin_frac', round(d['stages']['bin_kernel_frac'],3))
"

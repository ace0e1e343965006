#!/bin/bash
# Dev tool: A/B of environment switches on one workload.  usage: tools/ab.sh WORKLOAD "ENV1=a ENV2=b" "ENV3=c" ...
wl=$1; shift
for envs in "$@"; do
  out=$(env $envs python bench.py --workload $wl --no-configs --no-cpu-baseline --no-dropin --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d.get('roofline',{})
print('%.4f ms/step  min %.4f | conv per-level %s | alone %.4f ms (%.3f)' % (d['ms_per_step'], d['timing']['ms_per_step_min'], ['%.4f'%x for x in r.get('per_level_ms',[])], r.get('alone',{}).get('ms_per_launch',0), r.get('alone',{}).get('frac',0)))")
  echo "[$envs] $out"
done

#!/bin/bash
# Forced plan shapes around the planner's choice at the depths of the last slice of an 8-way split (one GPU): is there a
# better shape than the one the cost model picks?  bash tools/gpu_shapes.sh TAG
set -u
mkdir -p gpurun_out
T=$1
for DEPTH in 9 12 15; do
  python tools/profile_launch.py --kind filter --depth $DEPTH --reps 3 2>&1 | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); p=d.get('plan',{})
print('depth $DEPTH free   ', '%.4f ms' % d['ms'], 'tp', p.get('tprime'), 'K', p.get('steps'), 'J', p.get('segments'), 'fbits', p.get('filter_bits'), 'model %.3f' % p.get('cost',0))
" | tee -a gpurun_out/${T}_shapes.txt
  for TP in 15 16 17 18; do for J in 1 2 4 8; do
    python tools/profile_launch.py --kind filter --depth $DEPTH --reps 3 --force 2,$TP,$J 2>&1 | python -c "
import json,sys
try:
    d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); p=d.get('plan',{})
    print('depth $DEPTH 2,$TP,$J', '%.4f ms' % d['ms'], 'tp', p.get('tprime'), 'K', p.get('steps'), 'J', p.get('segments'), 'fbits', p.get('filter_bits'), 'model %.3f' % p.get('cost',0), 'launches', d.get('launches'))
except Exception as e:
    print('depth $DEPTH 2,$TP,$J failed', e)
" | tee -a gpurun_out/${T}_shapes.txt
  done; done
done

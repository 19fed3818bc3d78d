#!/bin/bash
# A/B of kernel build variants (qimcifa_b200/build.py build_variant) on one GPU: the bench launch shape at three depths and
# three sizes, best of 5 launches each.  bash tools/gpu_ab.sh TAG [variant ...]   ("" = the default library)
set -u
mkdir -p gpurun_out
T=$1; shift
for V in "$@"; do
  export QCF_LIB_VARIANT=$V
  [ "$V" = "default" ] && unset QCF_LIB_VARIANT
  for ARGS in "--bits 96" "--bits 96 --depth 8" "--bits 96 --depth 12" "--bits 96 --depth 64" "--bits 112" "--bits 128"; do
    python tools/profile_launch.py --kind filter $ARGS --reps 5 2>&1 | python -c "
import json,sys
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('$V', '$ARGS', '%.4f ms' % d['ms'], '%.4e g/s' % (d['guesses']/d['ms']*1e3), d.get('plan',{}).get('degree'), d.get('plan',{}).get('steps'), d.get('plan',{}).get('segments'), d.get('plan',{}).get('filter_bits'))
" | tee -a gpurun_out/${T}_ab.txt
  done
  python tools/slice_profile.py 7 8 448 25 2>&1 | awk -v v="$V" '{s+=$3; n+=1} END {printf "%s slice 7 of 8: %d steps of 448 batches, %.2f ms in all\n", v, n, s}' | tee -a gpurun_out/${T}_ab.txt
  python tools/slice_profile.py 7 8 896 25 2>&1 | awk -v v="$V" '{s+=$3; n+=1} END {printf "%s slice 7 of 8: %d steps of 896 batches, %.2f ms in all\n", v, n, s}' | tee -a gpurun_out/${T}_ab.txt
done

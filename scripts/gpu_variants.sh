#!/bin/bash
# Tuning: build variants ON THE BOX with different -D flags and bench each.
# usage: gpu_variants.sh "name1:DEF=1,DEF2=2" "name2:..." -- "bench args A" "bench args B" ...
set -u
mkdir -p gpurun_out
VARS=(); while [ $# -gt 0 ] && [ "$1" != "--" ]; do VARS+=("$1"); shift; done; shift
for v in "${VARS[@]}"; do
  name=${v%%:*}; defs=${v#*:}
  if [ "$name" != "base" ]; then
    if [ ! -f rust-swift-raytracer_b200/lib_$name/libraytracer.so ]; then
      RT_BUILD_VARIANT=$name RT_BUILD_DEFINES=$defs python rust-swift-raytracer_b200/build.py > gpurun_out/build_$name.log 2>&1 || { echo "build $name failed"; tail -5 gpurun_out/build_$name.log; continue; }
    fi
    grep -A2 "rt_render_kernelILb[01]ELb1ELi256" rust-swift-raytracer_b200/lib_$name/build.log | grep -i "registers" | head -2
    export RT_LIB_VARIANT=$name
  else
    unset RT_LIB_VARIANT
  fi
  for args in "$@"; do
    python bench.py $args --no-cpu-baseline --no-e2e 2>>gpurun_out/bench.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']
    print('$name', '|', '$args', '|', d['config']['workload'][:3], d['config']['kernel'][:5], 'ms/step %.3f' % d['ms_per_step'], 'Mrays/s %.1f' % d['value'], 'frac %.4f' % r['frac'])
"
  done
done

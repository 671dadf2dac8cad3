#!/bin/bash
# tuning helper (GPU box): run bench.py kernel-only under several env variants, from the repo root:
#   tools/variants.sh default relmajor ...
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/var_$name.json'))
    p=d['config']['passes']
    print('$name', 'ms/step=%.3f'%d['ms_per_step'], ' '.join('%s=%.2f'%(k,v['avg_ms']) for k,v in p.items() if v['avg_ms']>0.05))
except Exception as e:
    print('$name FAILED', e, open('gpurun_out/var_$name.err').read()[-400:])
PY
}
for v in "$@"; do
  case $v in
    default) run default A=1;;
    ring) run ring RGCN_B200_ETILE=0;;
    relmajor) run relmajor RGCN_B200_RANGE_NODES=1000000000;;
    nr4k) run nr4k RGCN_B200_RANGE_NODES=4096;;
    nr64k) run nr64k RGCN_B200_RANGE_NODES=65536;;
  esac
done
for v in "$@"; do
  case $v in
    t8) run t8 RGCN_B200_SPLIT=8;;
    t16) run t16 RGCN_B200_SPLIT=16;;
    t64) run t64 RGCN_B200_SPLIT=64;;
    t16c64) run t16c64 RGCN_B200_SPLIT=16 RGCN_B200_CHUNK=64;;
    t32c64) run t32c64 RGCN_B200_SPLIT=32 RGCN_B200_CHUNK=64;;
  esac
done

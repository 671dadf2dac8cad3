#!/bin/bash
# tuning helper (GPU box): run bench.py kernel-only under several env variants, from the repo root:
#   tools/variants.sh name1:ENV=V,ENV2=V name2: ...
mkdir -p gpurun_out
STEPS=${STEPS:-10}
run() { name=$1; shift; env "$@" python bench.py --steps $STEPS --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/var_$name.json'))
    p=d['config']['passes']
    print('$name', 'ms/step=%.3f'%d['ms_per_step'], ' '.join('%s=%.3f'%(k,v['avg_ms']) for k,v in p.items() if v['avg_ms']>0.03), flush=True)
except Exception as e:
    print('$name FAILED', e, open('gpurun_out/var_$name.err').read()[-600:], flush=True)
PY
}
for spec in "$@"; do
  name=${spec%%:*}; envs=${spec#*:}
  run "$name" A=1 ${envs//,/ }
done

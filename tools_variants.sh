#!/bin/bash
# tuning helper (GPU box): run bench.py kernel-only under several build/env variants
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
run default A=1
run breg0 RGCN_B200_BREG=0
run relmajor RGCN_B200_RANGE_NODES=1000000000
run relmajor_breg0 RGCN_B200_RANGE_NODES=1000000000 RGCN_B200_BREG=0
run nr16k RGCN_B200_RANGE_NODES=16384
run nr16k_breg0 RGCN_B200_RANGE_NODES=16384 RGCN_B200_BREG=0
run nr1k RGCN_B200_RANGE_NODES=1024

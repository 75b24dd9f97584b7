#!/bin/bash
# usage: profiles/step_ab.sh label [env assignments...]  -> one line with env-steps/s, ms/step, frac
label=$1; shift
env "$@" python bench.py --steps 200 --warmup 20 --skip-cpu-baseline --skip-e2e --skip-ppo --no-sweep 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', '%.4e' % d['value'], '%.4f ms' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'])"

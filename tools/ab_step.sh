#!/bin/bash
# step time of bench.py under a list of environment settings (one JSON line each): tools/ab_step.sh "A=1" "B=2 C=3" ...
for cfg in "$@"; do
  env $cfg timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$cfg', round(d['ms_per_step'],4), d['gpu_launches'])"
done

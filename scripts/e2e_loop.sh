#!/bin/bash
# Run ON THE GPU BOX: ten back-to-back default bench runs, printing ms_per_step / e2e ms_per_step / SM clock of each
# (stability of the end-to-end figure: before the streaming warm-up covered all pipeline slots, one run in ten read +40 %).
for i in 1 2 3 4 5 6 7 8 9 10; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3), d['clocks']['sm_mhz'])"; done

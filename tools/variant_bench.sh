#!/usr/bin/env bash
# time every tuning build in build/variants with the quick device-timed bench
for f in build/variants/libqg_*.so; do
  r=$(QG_LIB=$PWD/$f python bench.py --profile --steps 30 --warmup 80 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.3f ms  %.3e steps/s' % (d['ms_per_step'], d['value']))")
  echo "$f  $r"
done

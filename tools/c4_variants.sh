# time every tuning build in build/variants on the C3, C4 and C2 workloads (device-timed)
for f in build/variants/libqg_*.so; do for w in c3 c4 c2; do
  r=$(QG_LIB=$PWD/$f python bench.py --workload $w --profile --steps 48 --warmup 24 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  %.3e' % (d['ms_per_step'], d['value']))")
  echo "$w $f: $r"
done; done

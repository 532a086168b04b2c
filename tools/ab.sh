# A/B of environment switches on the GPU box: bash tools/ab.sh <tag> "ENV1=a ENV2=b" "ENV1=c" ...   ("-" = no switches)
# -> gpurun_out/<tag>_bench_<i>.log, gpurun_out/<tag>_ops_<i>.json and one summary line per configuration
T=$1; shift
mkdir -p gpurun_out
i=0
for cfg in "$@"; do
  [ "$cfg" = "-" ] && cfg=""
  env $cfg timeout 600 python bench.py --steps ${STEPS:-60} --latency-iters ${LAT:-200} --no-cpu-baseline ${BENCH_ARGS:-} > gpurun_out/${T}_bench_$i.log 2>&1; rc=$?
  cp gpurun_out/ops_profile.json gpurun_out/${T}_ops_$i.json 2>/dev/null
  python - <<PY
import json
try:
    l=[x for x in open('gpurun_out/${T}_bench_$i.log') if x.startswith('{')][-1]
    d=json.loads(l)
    print('$T [$cfg] rc=$rc value',round(d['value']),'serial',round(d.get('value_serial',0)),'e2e',round(d['e2e']['value']),'lat p50',round(d['latency']['p50_ms'],4),'sum_launch',round(d['roofline']['sum_launch_ms'],3),'top',d['roofline']['kernel'],round(d['roofline']['frac'],3))
except Exception as e:
    print('$T [$cfg] rc=$rc FAILED', e)
PY
  i=$((i+1))
done

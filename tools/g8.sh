# One 8-GPU box visit: the bare host->device ceiling at 1 / 2 / 4 / 8 concurrent ranks (tools/probe_h2d.py) and the 8-GPU bench.
mkdir -p gpurun_out
T=${1:-r2}
timeout 120 python tools/probe_h2d.py > gpurun_out/${T}_h2d_1.jsonl 2> gpurun_out/${T}_h2d_1.err
for n in 2 4 8; do
  timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 tools/probe_h2d.py > gpurun_out/${T}_h2d_$n.jsonl 2> gpurun_out/${T}_h2d_$n.err
done
cat gpurun_out/${T}_h2d_*.jsonl | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['n_gpus'], d['memory'], 'affinity', d['cpu_affinity'], 'per-rank min %.1f mean %.1f GB/s, aggregate %.1f GB/s -> %.0f frames/s' % (d['per_rank_gbs_min'], d['per_rank_gbs_mean'], d['aggregate_gbs'], d['frames_per_s_ceiling']))
"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 60 --warmup 5 --latency-iters 0 --no-cpu-baseline > gpurun_out/${T}_bench_8gpu.log 2>&1; echo "bench8 exit $?"
grep '^{' gpurun_out/${T}_bench_8gpu.log | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('8 GPUs: value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', d['ms_per_step'])
"

mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 300 python tools/layer_sweep.py --out gpurun_out/r2a_layer_sweep_f16silu.md > gpurun_out/r2a_sweep1.log 2>&1; tail -8 gpurun_out/r2a_sweep1.log
XRSEG_LIB_VARIANT=silu32 timeout 300 python tools/layer_sweep.py --out gpurun_out/r2a_layer_sweep_f32silu.md > gpurun_out/r2a_sweep2.log 2>&1; tail -8 gpurun_out/r2a_sweep2.log
timeout 900 python bench.py > gpurun_out/r2a_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r2a_bench.log
tail -c 3000 gpurun_out/r2a_bench.log
cp gpurun_out/ops_profile.json gpurun_out/r2a_ops_profile.json
XRSEG_LIB_VARIANT=silu32 timeout 600 python bench.py --steps 50 --latency-iters 0 --no-cpu-baseline > gpurun_out/r2a_bench_silu32.log 2>&1; echo "bench exit $?" >> gpurun_out/r2a_bench_silu32.log
tail -c 1200 gpurun_out/r2a_bench_silu32.log
cp gpurun_out/ops_profile.json gpurun_out/r2a_ops_profile_silu32.json

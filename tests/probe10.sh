mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p10_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/p10_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p10_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/p10_bench.log
tail -4 gpurun_out/p10_pytest.log; head -c 700 gpurun_out/p10_bench.log

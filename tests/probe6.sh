mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "tcgen05" > gpurun_out/p6_conv.log 2>&1; echo "conv exit $?" >> gpurun_out/p6_conv.log
tail -5 gpurun_out/p6_conv.log
timeout 900 python -m pytest tests -m gpu -x -q -k "not tcgen05" > gpurun_out/p6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/p6_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p6_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/p6_bench.log
tail -3 gpurun_out/p6_pytest.log; tail -c 1800 gpurun_out/p6_bench.log

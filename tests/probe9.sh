mkdir -p gpurun_out
timeout 400 python tests/probe_skip.py > gpurun_out/p_skip.log 2>&1; grep -E "xrseg_debug_conv|cycles per|Error|error" gpurun_out/p_skip.log | grep -A8 "skip 0" | head -40
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p9_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/p9_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/p9_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/p9_bench.log
tail -3 gpurun_out/p9_pytest.log; tail -c 1500 gpurun_out/p9_bench.log

mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/p2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/p2_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/p2_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/p2_smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/p2_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/p2_bench.log
tail -3 gpurun_out/p2_pytest.log; tail -2 gpurun_out/p2_smoke.log; tail -c 3000 gpurun_out/p2_bench.log

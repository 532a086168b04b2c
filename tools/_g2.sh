mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_pytest.log
tail -40 gpurun_out/r2b_pytest.log
timeout 300 python tools/layer_sweep.py --out gpurun_out/r2b_layer_sweep.md --frames coco139,coco632,coco2006,coco4495,coco7108,bus > gpurun_out/r2b_sweep.log 2>&1; tail -8 gpurun_out/r2b_sweep.log

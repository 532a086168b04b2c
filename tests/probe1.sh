mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/p1_smi.txt 2>&1
for s in "conv_direct" "conv_umma 1" "conv_umma 3" "conv_umma 0" "conv_umma 2" "pipeline direct" "pipeline umma" "layers"; do
  echo "=== $s" >> gpurun_out/p1.log
  timeout 150 python tests/gpu_probe.py $s >> gpurun_out/p1.log 2>&1
  echo "exit $?" >> gpurun_out/p1.log
done
tail -5 gpurun_out/p1.log

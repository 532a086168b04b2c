mkdir -p gpurun_out
python tests/ncu_skip.py > gpurun_out/p8_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_halo -c 2 -o gpurun_out/r1_conv_skip python tests/ncu_skip.py > gpurun_out/p8_ncu.log 2>&1
echo "ncu exit $?"

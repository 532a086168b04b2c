mkdir -p gpurun_out
python tests/ncu_convs.py > gpurun_out/p5_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_umma -c 5 -o gpurun_out/r1_conv_v1 python tests/ncu_convs.py > gpurun_out/p5_ncu.log 2>&1
echo "ncu exit $?"

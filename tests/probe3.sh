mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --profile-ops 0 --no-cpu-baseline"
$CMD > gpurun_out/p3_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches_v0.csv $CMD > gpurun_out/p3_ncu1.log 2>&1
echo "ncu1 exit $?"
python tests/ncu_convs.py > gpurun_out/p3_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_umma -c 5 -o gpurun_out/r1_conv_v0 python tests/ncu_convs.py > gpurun_out/p3_ncu2.log 2>&1
echo "ncu2 exit $?"
ls -la gpurun_out

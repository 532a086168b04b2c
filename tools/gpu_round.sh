# One GPU-box visit: parity tests, bench (both arms), ncu launch list, a section capture of one whole pass (CSV) and
# full captures (with source) of the dominant kernels.   usage: bash tools/gpu_round.sh <tag>   -> gpurun_out/<tag>_*
T=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/${T}_bench.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/${T}_bench_ref.log
cp gpurun_out/ops_profile.json gpurun_out/${T}_ops_profile.json 2>/dev/null
tail -3 gpurun_out/${T}_pytest.log; tail -c 1500 gpurun_out/${T}_bench.log; tail -c 600 gpurun_out/${T}_bench_ref.log
[ "${NCU:-1}" = "1" ] || exit 0
N=$(timeout 600 python tools/ncu_step.py 2> gpurun_out/${T}_ncu_plain.err | tee gpurun_out/${T}_ncu_plain.log | sed -n 's/^launches per pass \([0-9]*\).*/\1/p')
echo "launches per pass: $N"
[ -n "$N" ] || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $N -c $N --csv --log-file gpurun_out/${T}_launches.csv python tools/ncu_step.py > gpurun_out/${T}_ncu1.log 2>&1; echo "ncu launches exit $?"
timeout 1200 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum,smsp__cycles_active.avg \
  --clock-control none -s $N -c $N --page raw --csv --log-file gpurun_out/${T}_step_sections.csv python tools/ncu_step.py > gpurun_out/${T}_ncu2.log 2>&1; echo "ncu sections exit $?"
full() {  # name regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/${T}_full_$1 python tools/ncu_step.py > gpurun_out/${T}_ncu_$1.log 2>&1; echo "ncu full $1 exit $?"
}
NC=$(grep -c conv_halo_tma gpurun_out/${T}_launches.csv)   # TMA conv launches per pass (launch order = op order)
echo "conv_halo_tma launches per pass: $NC"
full stem stem_ 1 1                             # stem_rows_kernel (packed RGB) or stem_mma_kernel
full s2_b1 conv_halo_tma $((NC + 0)) 1         # b1 (3x3 stride 2, parity-plane TMA) of pass 2
full flat_b2cv1 conv_halo_tma $((NC + 1)) 1    # b2.cv1 (1x1, flat TMA) of pass 2
full halo_protocv2 conv_halo_tma $((NC + NC - 1)) 1  # proto.cv2+cv3 (3x3 stride 1, halo TMA, fused 1x1: the last conv) of pass 2
full bneck_b2m0 bottleneck_mma 3 1             # b2.m0 (fused Bottleneck) of pass 2
ND=$(grep -c dwconv3x3 gpurun_out/${T}_launches.csv)   # depthwise launches per pass (6 since the C2PSA pe rides in the attention kernel)
full dw dwconv3x3 $((ND + 1)) 1                # h3.cls.1dw of pass 2
full decode decode_filter 1 1
full maskprob mask_prob 1 1
full attention attention_kernel 1 1            # C2PSA attention + fused positional encoding
ls -la gpurun_out/${T}_*.ncu-rep

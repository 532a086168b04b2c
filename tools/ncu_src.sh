# Source-level ncu captures (--set full --import-source on) of single TMA conv launches of pass 2 of tools/ncu_step.py.
# usage: bash tools/ncu_src.sh <tag> <name>:<index among the conv_halo_tma launches of a pass> ...   -> gpurun_out/<tag>_src_<name>.ncu-rep
T=$1; shift
NC=${NC:-91}    # conv_halo_tma launches per pass (tools/gpu_round.sh prints it)
for spec in "$@"; do
  name=${spec%%:*}; idx=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_halo_tma -s $((NC + idx)) -c 1 -f -o gpurun_out/${T}_src_${name} python tools/ncu_step.py > gpurun_out/${T}_ncu_${name}.log 2>&1; echo "ncu $name exit $?"
done

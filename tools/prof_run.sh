#!/bin/bash
# GPU-box profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then --set full
# captures of one steady-state launch of every hot kernel.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-prof}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-gen --no-cpu"
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"k_post_fwd_umma|k_post_bwd_umma|k_wgrad_umma" -s 5 -c 5 -o gpurun_out/${TAG}_post $CMD > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu post exit $?"
ncu --set full --clock-control none --import-source on -k regex:"k_layer_fwd_p_umma|k_layer_bwd_gate_umma|k_layer_bwd_dx_p_umma" -s 195 -c 9 -o gpurun_out/${TAG}_layer $CMD > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu layer exit $?"
ls -la gpurun_out/${TAG}_*

#!/bin/bash
# GPU-box profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then --set full
# captures of one steady-state launch of every hot kernel.  Outputs under gpurun_out/<tag>_*.
TAG=${1:-prof}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-gen --no-cpu"
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
python bench.py --workload wide --steps 20 --no-gen --no-cpu > gpurun_out/${TAG}_bench_wide.json 2> gpurun_out/${TAG}_bench_wide.err; echo "bench wide exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; echo "bench reference exit $?"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1; echo "ncu list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"k_post_fwd_umma|k_post_bwd_umma|k_wgrad_umma" -s 5 -c 5 -o gpurun_out/${TAG}_post $CMD > gpurun_out/${TAG}_ncu2.log 2>&1; echo "ncu post exit $?"
# layer kernels: 30 forward + 30 backward launches per step; skip two steps, then 4 forward + 4 backward of the middle
ncu --set full --clock-control none --import-source on -k regex:"k_layer_fwd_p_umma" -s 73 -c 3 -o gpurun_out/${TAG}_lfwd $CMD > gpurun_out/${TAG}_ncu3.log 2>&1; echo "ncu layer fwd exit $?"
ncu --set full --clock-control none --import-source on -k regex:"k_layer_bwd_fused_umma" -s 73 -c 3 -o gpurun_out/${TAG}_lbwd $CMD > gpurun_out/${TAG}_ncu4.log 2>&1; echo "ncu layer bwd exit $?"
# generator: the benchmarked kernel (3x10, 256 streams) and the reference's own arch1 shape (S = P = 512 + GC, 10 streams)
ncu --set full --clock-control none --import-source on -k regex:"k_gen2" -s 1 -c 1 -o gpurun_out/${TAG}_gen python tools/gen_time.py 256 400 > gpurun_out/${TAG}_ncu5.log 2>&1; echo "ncu gen exit $?"
ls -la gpurun_out/${TAG}_*

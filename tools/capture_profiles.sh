#!/bin/bash
# Round profile capture (run under gpurun, one GPU):  bash tools/capture_profiles.sh
# 1. launch list of the bench command (cold-cache serialised times; shares must agree with bench.py's events)
# 2. full-set capture of one 32-clip step (second pass of tools/ncu_target.py)
set -x
mkdir -p gpurun_out
if [ -z "$SKIP_LAUNCH_LIST" ]; then
python bench.py --steps 2 --warmup 3 --no-extra --no-gpu-baseline --no-cpu-baseline --no-e2e --no-parity > gpurun_out/cap_bench.json 2> gpurun_out/cap_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-extra --no-gpu-baseline --no-cpu-baseline --no-e2e --no-parity > gpurun_out/ncu_launch.log 2>&1
fi
python tools/ncu_target.py 32 > gpurun_out/target.log 2>&1 || exit 1
TOT=$(grep -o 'launches [0-9]*' gpurun_out/target.log | awk '{print $2}')
NL=$((TOT / 2)); SKIP=$((TOT - NL))        # the first pass may carry one-time launches
ncu --set full --clock-control none --import-source on -s $SKIP -c $NL -f -o gpurun_out/step_b32 python tools/ncu_target.py 32 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/step_b32.ncu-rep --page raw --csv > gpurun_out/step_b32_raw.csv
rm -f gpurun_out/step_b32.ncu-rep        # ~100 MB: over gpurun's 64 MiB return limit; the raw page is what gets summarised
ls -la gpurun_out
AFB200_TRACE=1 python tools/trace_layers.py 32 2> gpurun_out/layer_trace_b32.raw > /dev/null
python tools/trace_summary.py gpurun_out/layer_trace_b32.raw 32 60 > gpurun_out/layer_trace_b32.txt
# K1 on its own: crop_kernel (af_crop_infer's feeder) and pack_u8_kernel (af_infer_u8's feeder), second iteration
ncu --set full --clock-control none -k regex:"crop_kernel|pack_u8_kernel" -s 2 -c 2 -f -o gpurun_out/k1 python tools/ncu_target_k1.py 32 > gpurun_out/ncu_k1.log 2>&1
ncu -i gpurun_out/k1.ncu-rep --page raw --csv > gpurun_out/k1_raw.csv
rm -f gpurun_out/k1.ncu-rep

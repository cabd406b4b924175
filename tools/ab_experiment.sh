# A/B runs of experiment builds (python -c "from neural_renderer_b200 import build; build.build_library(defines=(...), out=...)")
set -x
for ts in 2 4; do
python tools/kernel_times.py --steps 10 --ts $ts > gpurun_out/ab_stage_ts$ts.json
NR_B200_LIB=$PWD/neural_renderer_b200/exp_libD.so python tools/kernel_times.py --steps 10 --ts $ts > gpurun_out/ab_direct_ts$ts.json
done
python tools/kernel_times.py --steps 5 --batch 8 --faces 1000000 --size 1024 --ts 2 > gpurun_out/ab_stage_1M.json
NR_B200_LIB=$PWD/neural_renderer_b200/exp_libD.so python tools/kernel_times.py --steps 5 --batch 8 --faces 1000000 --size 1024 --ts 2 > gpurun_out/ab_direct_1M.json

# A/B runs of experiment builds (python -c "from neural_renderer_b200 import build; build.build_library(defines=(...), out=...)")
set -x
python tools/kernel_times.py --steps 10 > gpurun_out/ab_fast.json
for v in nofast cheap; do
NR_B200_LIB=$PWD/neural_renderer_b200/exp_$v.so python tools/kernel_times.py --steps 10 > gpurun_out/ab_$v.json
done
python tools/kernel_times.py --steps 5 --batch 8 --faces 4928 --size 512 --aa 1 > gpurun_out/ab_fast_aa.json
NR_B200_LIB=$PWD/neural_renderer_b200/exp_nofast.so python tools/kernel_times.py --steps 5 --batch 8 --faces 4928 --size 512 --aa 1 > gpurun_out/ab_nofast_aa.json
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3
NR_B200_LIB=$PWD/neural_renderer_b200/exp_cheap.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3

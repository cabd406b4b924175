# A/B runs of experiment builds (python -c "from neural_renderer_b200 import build; build.build_library(defines=(...), out=...)")
set -x
python tools/kernel_times.py --steps 10 > gpurun_out/ab_base.json
for v in f6 f8 e9 t8; do
NR_B200_LIB=$PWD/neural_renderer_b200/exp_$v.so python tools/kernel_times.py --steps 10 > gpurun_out/ab_$v.json
done

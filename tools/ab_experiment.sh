#!/bin/bash
# A/B runs of experiment builds on the GPU box (one gpurun call measures every variant on the SAME box).
#
#   here:    python -c "from neural_renderer_b200 import build; \
#                build.build_library(defines=('NR_ES_UNROLL=4',), out='neural_renderer_b200/exp_u4.so')"
#   then:    gpurun --timeout 900 -- 'bash tools/ab_experiment.sh u4 [more variants ...]'
#
# For every variant V the library neural_renderer_b200/exp_V.so is selected with NR_B200_LIB; the default build runs
# first and last (box drift shows as the difference between the two).  Results: gpurun_out/ab_<variant>.json (step time,
# e2e) and gpurun_out/kt_<variant>.json (per-kernel CUDA-event times).  The compile-time knobs that exist are listed at
# the top of csrc/nr_forward.cu and csrc/nr_backward.cu (NR_ES_UNROLL, NR_TG_COMBINE, NR_COL_SMEM_MAX_S, NR_BIG_AREA, ...).
set -x
mkdir -p gpurun_out
BENCH="python bench.py --no-side-measurements --no-shared-mesh --steps 50 --warmup 5"
run() {  # name, library ('' = default build)
    if [ -n "$2" ]; then export NR_B200_LIB=$2; else unset NR_B200_LIB; fi
    $BENCH > gpurun_out/ab_$1.json 2> gpurun_out/ab_$1.err
    python tools/kernel_times.py --steps 10 > gpurun_out/kt_$1.json 2>&1
}
run base ""
for v in "$@"; do run "$v" "$PWD/neural_renderer_b200/exp_$v.so"; done
run base2 ""
unset NR_B200_LIB
python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -x -q -m gpu 2>&1 | tail -3 > gpurun_out/ab_tests.txt

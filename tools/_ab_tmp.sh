set -x
python -m pytest tests -x -q -m gpu 2>&1 | tail -4 > gpurun_out/gputests.txt
B="python bench.py --no-side-measurements --no-shared-mesh --steps 50 --warmup 5"
$B > gpurun_out/ab_new.json 2>gpurun_out/ab_new.err
for v in nopad old; do
NR_B200_LIB=$PWD/neural_renderer_b200/exp_$v.so $B > gpurun_out/ab_$v.json 2>gpurun_out/ab_$v.err
done
for v in nopad; do
NR_B200_LIB=$PWD/neural_renderer_b200/exp_$v.so python tools/kernel_times.py --steps 10 > gpurun_out/kt_$v.json 2>&1
done
python tools/kernel_times.py --steps 10 > gpurun_out/kt_new.json 2>&1
python tools/kernel_times.py --steps 10 --ts 2 > gpurun_out/kt_new_ts2.json 2>&1
NR_B200_LIB=$PWD/neural_renderer_b200/exp_nopad.so python tools/kernel_times.py --steps 10 --ts 2 > gpurun_out/kt_nopad_ts2.json 2>&1

set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1f.log 2> gpurun_out/bench_r1f.err
python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/plain_f.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_f.log 2>&1
for k in blend_backward_material_kernel blend_forward_kernel deferred_backward_kernel deferred_shade_kernel preprocess_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/r1f_$k python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_full_$k.log 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:rs_onesweep_kernel -s 44 -c 2 -f -o gpurun_out/r1f_rs_onesweep_6bit python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_full_rs6.log 2>&1
ls gpurun_out | grep r1f

set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1e.log 2> gpurun_out/bench_r1e.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1e_ref.log 2> gpurun_out/bench_r1e_ref.err
python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/plain_e.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1e.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_e.log 2>&1
for k in rs_onesweep_kernel blend_backward_kernel blend_forward_kernel deferred_backward_kernel deferred_shade_kernel preprocess_kernel emit_keys_kernel deferred_loss_kernel geometry_chain_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/r1e_$k python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_full_$k.log 2>&1
done
ls gpurun_out | tail -5

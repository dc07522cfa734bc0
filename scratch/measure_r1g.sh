set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_g.log 2>&1; tail -2 gpurun_out/smoke_g.log
python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1g.log 2> gpurun_out/bench_r1g.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1g_ref.log 2> gpurun_out/bench_r1g_ref.err
python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/plain_g.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_g.log 2>&1
for k in blend_backward_material_kernel blend_forward_kernel deferred_backward_kernel deferred_loss_kernel preprocess_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/r1g_$k python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_full_$k.log 2>&1
done
python scratch/light_once.py > gpurun_out/light_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:cm_sparse -s 2 -c 2 -f -o gpurun_out/r1g_cm_sparse_kernel python scratch/light_once.py > gpurun_out/ncu_light.log 2>&1
ls gpurun_out | grep r1g

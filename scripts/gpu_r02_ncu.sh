# r02: ncu evidence of the shipped build (run via gpurun): launch list of the default bench, full captures of the four kernels
tag=$1
mkdir -p gpurun_out
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch_$tag.log 2>&1
python scripts/quick_perf.py c2 > gpurun_out/quick_c2_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_price_batch -s 3 -c 1 -f -o gpurun_out/prof_k_price_$tag \
    python scripts/quick_perf.py c2 > gpurun_out/ncu_c2_$tag.log 2>&1
python scripts/quick_perf.py c3 > gpurun_out/quick_c3_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_price_dense -s 3 -c 1 -f -o gpurun_out/prof_k_dense_$tag \
    python scripts/quick_perf.py c3 > gpurun_out/ncu_c3_$tag.log 2>&1
python scripts/ncu_targets.py loss > gpurun_out/quick_loss_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_loss_batch -s 3 -c 1 -f -o gpurun_out/prof_k_loss_$tag \
    python scripts/ncu_targets.py loss > gpurun_out/ncu_loss_$tag.log 2>&1
python scripts/ncu_targets.py gen > gpurun_out/quick_gen_$tag.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_gen_market -s 1 -c 1 -f -o gpurun_out/prof_k_gen_market_$tag \
    python scripts/ncu_targets.py gen > gpurun_out/ncu_gen_$tag.log 2>&1
cat gpurun_out/quick_c2_$tag.log gpurun_out/quick_c3_$tag.log gpurun_out/quick_loss_$tag.log gpurun_out/quick_gen_$tag.log
tail -2 gpurun_out/ncu_gen_$tag.log

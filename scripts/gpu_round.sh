# Round GPU session: parity tests, smoke, bench, ncu launch list + one full capture (run via gpurun).
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q -s -x 2>&1 | tail -30
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; tail -2 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
python bench.py --steps 5 --warmup 3 --workload c3 --no-cpu-baseline > gpurun_out/bench_c3.json 2>> gpurun_out/bench_c2.err; cat gpurun_out/bench_c3.json
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_price_batch -s 1 -c 1 -o gpurun_out/prof_k_price_r01 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log

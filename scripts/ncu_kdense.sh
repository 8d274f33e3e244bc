# One full ncu capture of k_price_dense on the C3 shape (developer probe; run via gpurun).
mkdir -p gpurun_out
python scripts/quick_perf.py c3 > gpurun_out/quick_dense_$1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_price_dense -s 3 -c 1 -f -o gpurun_out/prof_k_dense_$1 \
    python scripts/quick_perf.py c3 > gpurun_out/ncu_dense_$1.log 2>&1
cat gpurun_out/quick_dense_$1.log; tail -2 gpurun_out/ncu_dense_$1.log

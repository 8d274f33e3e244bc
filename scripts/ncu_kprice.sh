# One full ncu capture of k_price on the C2 shape (developer probe; run via gpurun).
# usage: bash scripts/ncu_kprice.sh <tag>
mkdir -p gpurun_out
python scripts/quick_perf.py c2 > gpurun_out/quick_$1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_price_batch -s 3 -c 1 -f -o gpurun_out/prof_k_price_$1 \
    python scripts/quick_perf.py c2 > gpurun_out/ncu_$1.log 2>&1
cat gpurun_out/quick_$1.log; tail -2 gpurun_out/ncu_$1.log

# One full ncu capture of k_loss_batch on the C5 round shape (developer probe; run via gpurun).
# usage: bash scripts/ncu_kloss.sh <tag>
mkdir -p gpurun_out
python scripts/quick_perf.py loss > gpurun_out/quick_loss_$1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_loss_batch -s 3 -c 1 -f -o gpurun_out/prof_k_loss_$1 \
    python scripts/quick_perf.py loss > gpurun_out/ncu_loss_$1.log 2>&1
cat gpurun_out/quick_loss_$1.log; tail -2 gpurun_out/ncu_loss_$1.log

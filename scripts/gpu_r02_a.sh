# r02 session A: baseline at HEAD of round 1 (tests, quick perf, three full ncu captures)
mkdir -p gpurun_out
set -x
python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python scripts/quick_perf.py all 2>&1 | tee gpurun_out/quick_r02_base.log
bash scripts/ncu_kprice.sh r02_base | tail -3
bash scripts/ncu_kdense.sh r02_base | tail -3
bash scripts/ncu_kloss.sh r02_base | tail -3

TAG=$1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"prb_setup|prb_pgs|prb_ik" -s 7300 -c 80 --csv --log-file gpurun_out/launches64k_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b_l64k_$TAG.log 2>&1
python tools/exp_usage.py 2>&1 | tail -8

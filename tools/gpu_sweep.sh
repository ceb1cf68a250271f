# A/B of the arm-island kernel's block size (envs per warp): PRB_ARM_THREADS = 4 x envs per block
for T in 8 16 32; do
  PRB_ARM_THREADS=$T python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/sweep_$T.json
done

bash scripts/gpu_run_a.sh
bash scripts/gpu_run_c.sh

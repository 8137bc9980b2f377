#!/bin/bash
# full GPU parity suite (as the driver runs it), then the bench line, smoke, and the ncu launch list of the same bench command
mkdir -p gpurun_out; rm -f gpurun_out/full_rc.txt
timeout 1500 python -m pytest tests/ -x -q -m gpu --timeout 600 > gpurun_out/full_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/full_rc.txt
tail -n 4 gpurun_out/full_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/full_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/full_rc.txt
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?" >> gpurun_out/full_rc.txt
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/full_bench_ref.json 2> gpurun_out/full_bench_ref.err; echo "ref rc=$?" >> gpurun_out/full_rc.txt
if [ -n "$FULL_NCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/full_launches.csv python bench.py --steps 2 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/full_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/full_rc.txt
fi
cat gpurun_out/full_rc.txt; head -c 600 gpurun_out/full_bench.json; echo; head -c 400 gpurun_out/full_bench_ref.json

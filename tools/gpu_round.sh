#!/bin/bash
# One gpurun call: parity tests, smoke, bench (both arms), ncu launch list + one full capture.
mkdir -p gpurun_out; rm -f gpurun_out/status.txt
nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "simt or kmeans or uniq or backward or empty" > gpurun_out/pytest_simt.log 2>&1; echo "pytest_simt rc=$?" >> gpurun_out/status.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_all.log 2>&1; echo "pytest_all rc=$?" >> gpurun_out/status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/status.txt
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/status.txt
timeout 300 python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?" >> gpurun_out/status.txt
timeout 300 python tools/profile_step.py > gpurun_out/prof_plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_list.log 2>&1
echo "ncu_list rc=$?" >> gpurun_out/status.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:rq_fwd_tc|rq_bwd" -c 14 -o gpurun_out/prof_tc -f python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "ncu_full rc=$?" >> gpurun_out/status.txt
cat gpurun_out/status.txt; tail -5 gpurun_out/pytest_all.log; cat gpurun_out/bench.json | head -c 3000

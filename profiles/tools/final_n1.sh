#!/bin/bash
# The N = 1 measurements kept under profiles/r2/ (bench line, reference arm, ncu launch list, spans / timelines of the
# large-batch and ensemble launches).   usage (GPU box): bash profiles/tools/final_n1.sh
set -x
python bench.py > gpurun_out/r2f_bench_n1.json 2> gpurun_out/r2f_bench_n1.err; echo "exit $?"
python bench.py --impl reference --steps 200 --warmup 5 > gpurun_out/r2f_bench_ref_n1.json 2> gpurun_out/r2f_bench_ref_n1.err; echo "exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 200 --csv --log-file gpurun_out/r2f_launches_per256.csv python bench.py --steps 60 --warmup 20 --no-extra --no-cpu > gpurun_out/r2f_ncu_bench.log 2>&1; echo "exit $?"
python profiles/tools/kernel_spans.py 65536 fp32 8 2>&1 | grep -v DEVICE > gpurun_out/r2f_spans_fp32_b65536_n1.txt
python profiles/tools/phase_timeline.py per256 65536 2>&1 | grep -v DEVICE | head -16 > gpurun_out/r2f_timeline_b65536_ws.txt
python profiles/tools/ensemble_timeline.py 2>&1 | grep -v DEVICE | head -16 > gpurun_out/r2f_ens_timeline.txt
tail -1 gpurun_out/r2f_bench_n1.json | cut -c1-600

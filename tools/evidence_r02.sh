#!/bin/bash
# Regenerates the round-2 evidence kept under profiles/ on a B200 box (one GPU):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/evidence_r02.sh'
# Every ncu pass runs only after the same program has exited 0 without ncu; no number printed under ncu is a bench value.
set -u
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/r02_gputest_final.log 2>&1; echo "pytest rc=$?" >> $O/r02_gputest_final.log
python bench.py > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_reference_arm.err
python tools/ref_golden_report.py > $O/r02_ref_golden_accuracy.txt 2>&1
python tools/accuracy_report.py > $O/r02_accuracy.txt 2>&1
# launch list of the bench command (per-launch times are cold-cache and serialised: the SHARES are what must agree)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-grad-leg --no-extra-legs > $O/r02_bench_for_ncu.json 2>/dev/null && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file $O/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-grad-leg --no-extra-legs > $O/r02_ncu_launches.log 2>&1
# full captures at the bench chunk size (65536 candidates per launch): forward call, value + gradient call
python tools/prof_run.py 65536 2 > $O/r02_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/prof_r02_fwd \
    python tools/prof_run.py 65536 2 > $O/r02_ncu_fwd.log 2>&1
python tools/prof_grad_run.py 65536 2 > $O/r02_prof_grad_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $O/prof_r02_grad \
    python tools/prof_grad_run.py 65536 2 > $O/r02_ncu_grad.log 2>&1
python tools/sgld_step_time.py > $O/r02_sgld_step_time.log 2>&1
python tools/sgld_loop_time.py > $O/r02_sgld_loop_time.log 2>&1
python tools/sgld_graph_step.py > $O/r02_sgld_graph_step.log 2>&1
python tools/sgld_loop_stages.py > $O/r02_sgld_loop_stages_after.log 2>&1
# the .ncu-rep files (60-70 MB each) exceed what gpurun copies back: summarise them here and drop them
{ echo "==== forward call, 65536 candidates (tools/prof_run.py) ===="; python tools/ncu_summary.py $O/prof_r02_fwd.ncu-rep;
  echo; echo "==== value + gradient call, 65536 candidates (tools/prof_grad_run.py) ===="; python tools/ncu_summary.py $O/prof_r02_grad.ncu-rep; } > $O/r02_ncu_summary_final.txt 2>&1
rm -f $O/prof_r02_fwd.ncu-rep $O/prof_r02_grad.ncu-rep
tail -3 $O/r02_gputest_final.log
head -c 400 $O/r02_bench_n1.json

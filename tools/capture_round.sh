#!/bin/bash
# One round-end capture on a B200 box (run through gpurun): GPU tests, both bench arms, per-kernel times, the ncu launch list of the bench
# command and one `ncu --set full` pass over the five kernels of a learn step.   usage: bash tools/capture_round.sh <tag>   -> gpurun_out/<tag>/
D=gpurun_out/${1:-capture}; mkdir -p $D
python -m pytest tests -q -m gpu -x > $D/gputests.log 2>&1; echo "tests rc=$?"; tail -2 $D/gputests.log
python bench.py --impl reference > $D/bench_reference.json 2> $D/bench_ref.err; echo "reference arm rc=$?"
python bench.py > $D/bench_n1.json 2> $D/bench.err; echo "bench rc=$?"
python tools/kernel_times.py bf16x3 100 > $D/kernel_times.txt 2>/dev/null; python tools/kernel_times.py bf16x3 200 >> $D/kernel_times.txt 2>/dev/null
python tools/kernel_times_match.py 32 > $D/kernel_times_match.txt 2>/dev/null; python tools/kernel_times_match.py 2 >> $D/kernel_times_match.txt 2>/dev/null
python tools/kernel_times_conv.py > $D/kernel_times_conv.txt 2>/dev/null
grep -h "avg=\|sum of" $D/kernel_times.txt | head -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $D/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $D/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'prep_all|inr_fwd_ts|learn_objective|inr_bwd_pp|learn_tail' -s 20 -c 5 -o $D/prof_step -f python tools/kernel_times.py > $D/ncu.log 2>&1
tail -2 $D/ncu.log

#!/bin/bash
# On the GPU box: bench (plain), then the ncu launch list of the same command, then one `--set full` capture of the
# step kernel and of the plain fit kernel.  Outputs under gpurun_out/ (tag = $1).
tag=${1:-r01b}
out=gpurun_out
python bench.py --steps 2 --warmup 1 > $out/${tag}_bench_short.log 2> $out/${tag}_bench_short.err || { echo "bench failed"; tail -5 $out/${tag}_bench_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 > $out/${tag}_ncu_bench.log 2>&1
MB_ONLY=step python tools/microbench.py || exit 1
MB_ONLY=step ncu --set full --clock-control none --import-source on -k regex:fit_kernel -s 4 -c 4 -f -o $out/${tag}_step \
    python tools/microbench.py > $out/${tag}_ncu_step.log 2>&1
ncu -i $out/${tag}_step.ncu-rep --page raw --csv > $out/${tag}_step_raw.csv
ncu -i $out/${tag}_step.ncu-rep --page source --csv > $out/${tag}_step_source.csv 2>/dev/null
ls -la $out | grep $tag

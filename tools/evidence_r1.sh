#!/bin/bash
# Round-1 evidence pass on the GPU box: full GPU test suite, benches of every workload, reference arm,
# ncu launch list + `--set full` capture of one apply (cfg3) and of the N_t = 16384 FFT kernel.
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_r1.log 2>&1; tail -2 $O/pytest_gpu_r1.log
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_r1_ref.json 2> $O/bench_r1_ref.err
timeout 400 python bench.py > $O/bench_r1_cfg3.json 2> $O/bench_r1_cfg3.err; tail -c 300 $O/bench_r1_cfg3.err
timeout 300 python bench.py --workload cfg4 --steps 10 > $O/bench_r1_cfg4.json 2> $O/bench_r1_cfg4.err
for w in cfg1 cfg2 cfg5; do timeout 200 python bench.py --workload $w --no-cpu > $O/bench_r1_$w.json 2> $O/bench_r1_$w.err; done
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-gmres"
timeout 200 $CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv \
    --log-file $O/r01_launches_cfg3.csv $CMD > $O/ncu_list.log 2>&1
timeout 200 $CMD > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pd_ -s 27 -c 9 -o $O/prof_r1_cfg3 $CMD > $O/ncu_full.log 2>&1
timeout 100 python tools/fft16k_ncu.py > $O/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pd_fft_16k -s 2 -c 2 -o $O/prof_r1_fft16k python tools/fft16k_ncu.py > $O/ncu_fft16k.log 2>&1
ls -la $O | tail -20

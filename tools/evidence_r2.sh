#!/bin/bash
# Round-2 evidence pass on ONE GPU: benches of every workload, reference arm, ncu launch list + `--set full`
# capture of one apply (cfg3).  Outputs under gpurun_out/ (copy what is to be judged into profiles/).
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/r02_pytest_gpu.log 2>&1; echo "rc=$?" >> $O/r02_pytest_gpu.log; tail -3 $O/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_cfg3.json 2> $O/r02_bench_cfg3.err; tail -c 300 $O/r02_bench_cfg3.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference_arm.json 2> $O/r02_bench_ref.err
for w in cfg1 cfg2 cfg5; do timeout 200 python bench.py --workload $w --no-cpu > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; done
[ -n "$EVIDENCE_SKIP_CFG4" ] || timeout 300 python bench.py --workload cfg4 --steps 10 > $O/r02_bench_cfg4.json 2> $O/r02_bench_cfg4.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-gmres --no-cfg4"
timeout 200 $CMD > $O/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file $O/r02_launches_cfg3.csv $CMD > $O/ncu_list.log 2>&1
timeout 200 $CMD > $O/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:pd_ -s 20 -c 5 -o $O/prof_r2_cfg3 $CMD > $O/ncu_full.log 2>&1
ls -la $O | tail -12

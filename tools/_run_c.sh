cd /root/repo
mkdir -p gpurun_out
for w in cfg1 cfg2; do
  PD_PDL=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:pd_ -c 60 --csv --log-file gpurun_out/r2_launches_${w}.csv python bench.py --workload $w --steps 3 --warmup 1 --no-cpu --no-gmres --no-cfg4 > gpurun_out/r2_ncu_${w}.log 2>&1
  echo "$w rc=$?"
done

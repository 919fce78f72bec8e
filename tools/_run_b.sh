cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_slab_peer.py tests/test_gpu_alpha.py -x -q -m gpu > gpurun_out/r2_pdl_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_pdl_tests.log
tail -5 gpurun_out/r2_pdl_tests.log
for w in cfg1 cfg2 cfg5 cfg3; do
  for p in 1 0; do
    PD_PDL=$p timeout 600 python bench.py --workload $w --steps 50 --warmup 5 --no-cpu --no-gmres --no-cfg4 > gpurun_out/r2_pdl_${w}_$p.json 2> gpurun_out/r2_pdl_${w}_$p.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/r2_pdl_${w}_$p.json").read().strip().splitlines()[-1])
print("$w pdl=$p", round(d["ms_per_step"],5), {k:round(v,4) for k,v in d.get("kernels_ms",{}).items()})
PY
  done
done

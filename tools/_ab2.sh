for i in 1 2; do
for v in 0 262144 65536; do
echo "OVERLAP_MAX_PIXELS=$v" >> gpurun_out/r2b_overlap_ab.txt
TVAE_WGRAD_OVERLAP_MAX_PIXELS=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2b_overlap_ab.txt
done
done
cat gpurun_out/r2b_overlap_ab.txt

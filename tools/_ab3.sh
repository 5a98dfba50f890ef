timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -k "fused_reconstruction or conv_fwd or stay_in_bounds" 2>&1 | tail -2
for i in 1 2 3; do
for v in 1 0; do
echo "FUSE_NLL=$v" >> gpurun_out/r2b_fuse_nll_ab.txt
TVAE_FUSE_NLL=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-secondary --no-cpu-baseline --skip-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'])" >> gpurun_out/r2b_fuse_nll_ab.txt
done
done
cat gpurun_out/r2b_fuse_nll_ab.txt

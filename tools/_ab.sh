set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "skinny or sub_block" 2>&1 | tail -3
PYTHONPATH=. timeout 120 python tools/wgrad_skinny_bench.py 2>&1 | tee gpurun_out/r2b_wgrad_skinny.txt
for i in 1 2; do
for v in 1 0; do
echo "SPLIT_WIDE_WGRAD=$v" >> gpurun_out/r2b_split_ab.txt
TVAE_SPLIT_WIDE_WGRAD=$v timeout 300 python tools/step_timeline.py 256 5 2>&1 | grep -E "ms/step live|1028|1024|skinny" >> gpurun_out/r2b_split_ab.txt
done
done
cat gpurun_out/r2b_split_ab.txt

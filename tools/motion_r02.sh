timeout 600 python -m pytest tests/test_gpu_operators.py tests/test_gpu_fuzz.py tests/test_gpu_graph.py tests/test_gpu_fullsize.py -q -m gpu -x 2>&1 | tail -6
timeout 200 python tools/kbench.py --ops motion61,motion61s,walk61 --batches 16,64 --iters 10 2>&1 | grep -v Warn | tee gpurun_out/r02l_motion.jsonl
echo rows-form; PSX_PSF_FORM=rows timeout 200 python tools/kbench.py --ops motion61s,walk61 --batches 16 --iters 10 2>&1 | grep -v Warn | tee gpurun_out/r02l_motion_rows.jsonl
echo scalar; PSX_NO_C2V2=1 timeout 200 python tools/kbench.py --ops motion61 --batches 16 --iters 10 2>&1 | grep -v Warn | tee gpurun_out/r02l_motion_scalar.jsonl

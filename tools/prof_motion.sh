ncu --set full --clock-control none --import-source on -k regex:'conv2d_rowseg2' --launch-skip 4 --launch-count 2 -o gpurun_out/r02m_motion -f python tools/kbench.py --ops motion61 --batches 16 --iters 1 > gpurun_out/r02m_ncu.log 2>&1
ncu -i gpurun_out/r02m_motion.ncu-rep --page raw --csv > gpurun_out/r02m_raw.csv 2>/dev/null
ncu -i gpurun_out/r02m_motion.ncu-rep --page details --csv > gpurun_out/r02m_details.csv 2>/dev/null
ls -la gpurun_out/r02m*

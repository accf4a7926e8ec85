set -x
python bench.py --steps 4 --warmup 3 --no-extra > gpurun_out/r02k_b.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r02k_ncu.log 2>&1
ONCE=1 python tools/mean_ab.py 16 > gpurun_out/r02k_once.log 2>&1 || exit 1
ONCE=1 ncu --set full --clock-control none --import-source on -k regex:'blur_k1_tc|k2_post' --launch-count 6 -o gpurun_out/r02k_prof -f python tools/mean_ab.py 16 > gpurun_out/r02k_ncu2.log 2>&1
ncu -i gpurun_out/r02k_prof.ncu-rep --page raw --csv > gpurun_out/r02k_raw.csv 2>/dev/null
ls -la gpurun_out | tail -5

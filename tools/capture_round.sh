set -x
python bench.py > gpurun_out/r02f_bench_n1.json 2> gpurun_out/r02f_bench_n1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_ref.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02f_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:oe_fbank2 -s 8 -c 2 -f -o gpurun_out/r02f_fbank2_instep python tools/step_breakdown.py > gpurun_out/r02f_ncu_full.log 2>&1
ncu -i gpurun_out/r02f_fbank2_instep.ncu-rep --page raw --csv > gpurun_out/r02f_raw.csv 2>/dev/null
ncu --metrics dram__bytes_read.sum --cache-control none --clock-control none -c 60 --csv --log-file gpurun_out/r02f_dram_read.csv python tools/step_breakdown.py > /dev/null 2>&1
ncu --metrics dram__bytes_write.sum --cache-control none --clock-control none -c 60 --csv --log-file gpurun_out/r02f_dram_write.csv python tools/step_breakdown.py > /dev/null 2>&1
tail -2 gpurun_out/r02f_ncu_full.log

#!/bin/bash
# Final captures of this session (run on the GPU box through gpurun): GPU tests, FLAC decode timings, N=1 bench line,
# reference arm, one ncu --set full capture of the FLAC decode kernel.
python -m pytest tests -m gpu -x -q > gpurun_out/s5_gputests.log 2>&1; tail -2 gpurun_out/s5_gputests.log
python tools/flac_gpu_bench.py > gpurun_out/s5_flac_bench.txt 2>&1; cat gpurun_out/s5_flac_bench.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/s5_bench_n1.json 2> gpurun_out/s5_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/s5_bench_ref.json 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:flac_decode -c 1 -s 8 -o gpurun_out/s5_flac_ncu python tools/flac_gpu_bench.py --reps 2 > gpurun_out/s5_flac_ncu.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/s5_bench_n1.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["from_flac_files"]["value"], d["e2e"]["from_flac_files"]["host_decode"]["value"],
      d["e2e"]["from_wav_files"]["value"], d["roofline"]["frac"])
r = json.load(open("gpurun_out/s5_bench_ref.json"))
print(r["value"], r.get("cpu_baseline"))
PY

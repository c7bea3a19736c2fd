#!/bin/bash
# Round-end evidence on one B200: GPU test suite, smoke, the default bench line + per-launch table, the ncu launch list of the
# bench command and one `ncu --set full` pass over the forward kernels (reports stay in /tmp; only csv/txt go to gpurun_out/).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/final_suite.txt; tail -2 gpurun_out/final_suite.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --layer-table gpurun_out/r02_layers_final2.csv > gpurun_out/r02_bench_final2.json 2> gpurun_out/r02_bench_final2.err
tail -c 600 gpurun_out/r02_bench_final2.json | head -c 300; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_bench_final.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-torch-baseline > gpurun_out/bench_under_ncu.json 2>/dev/null
SCENES=16 timeout 900 ncu --set full --clock-control none -k regex:"conv_tc|fusion_kernel|bev_pack" -c 60 -o /tmp/fwd python tools/prof_forward.py > /dev/null 2>&1
ncu -i /tmp/fwd.ncu-rep --page raw --csv > gpurun_out/r02_fwd_full_raw.csv 2>/dev/null
wc -l gpurun_out/r02_launches_bench_final.csv gpurun_out/r02_fwd_full_raw.csv

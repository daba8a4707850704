#!/bin/bash
# Round-2 measurement pass on ONE B200 (run under gpurun): bench lines for cfg2/cfg3/cfg4, the ncu launch list of a bench
# run, and `ncu --set full` captures of the top kernels.  Everything lands in gpurun_out/ (summaries are made on the CPU box
# with tools/summarize_profiles.py and committed under profiles/).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
R=${1:-r2}
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/${R}_gpu_info.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_cfg2_n1.json 2> gpurun_out/${R}_bench_cfg2_n1.err
python bench.py --config cfg3 --steps 6 --warmup 3 --no-eager-baseline > gpurun_out/${R}_bench_cfg3_n1.json 2> gpurun_out/${R}_bench_cfg3_n1.err
python bench.py --config cfg4 --steps 4 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/${R}_bench_cfg4_n1.json 2> gpurun_out/${R}_bench_cfg4_n1.err
BENCH="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/${R}_launches.csv $BENCH > gpurun_out/${R}_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 36 -c 4 -f -o gpurun_out/${R}_attn_full $BENCH > gpurun_out/${R}_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 100 -c 8 -f -o gpurun_out/${R}_gemm_full $BENCH > gpurun_out/${R}_ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gp_sample -c 1 -f -o gpurun_out/${R}_gp_full python tools/run_gp_once.py 512 > gpurun_out/${R}_ncu_gp.log 2>&1
cuobjdump -sass transformerscandobayesianinference_b200/libpfn_b200.so | grep -oE "UTCHMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|SYNCS|HMMA|LDGSTS|UBLKCP|FFMA2|FMUL2|FADD2" | sort | uniq -c > gpurun_out/${R}_sass_mnemonics.txt
for f in cfg2 cfg3 cfg4; do echo "== $f: $(python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${R}_bench_${f}_n1.json"))
    print(round(d["value"], 1), "seq/s", round(d["ms_per_step"], 2), "ms/step; e2e", round(d["e2e"]["value"], 1), "; step frac", round(d["roofline"]["step"]["frac"], 4), "gemm frac", round(d["roofline"]["frac"], 4), d["clocks"])
    if d.get("gpu_eager_baseline"): print(d["gpu_eager_baseline"])
    if d.get("cpu_baseline"): print(d["cpu_baseline"])
except Exception as e:
    print("FAILED", e)
PY
)"; done
tail -3 gpurun_out/${R}_bench_cfg4_n1.err

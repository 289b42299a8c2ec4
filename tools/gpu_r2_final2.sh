#!/bin/bash
# round 2, final refresh on one B200 after the last kernel changes: test-suite, smoke, the driver's bench lines, launch lists,
# full ncu capture of the GEMMs (the larger sweep of workloads and the CUDA-core capture: tools/gpu_r2_final.sh)
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8 | cut -c1-300 | tee $O/r02_gpu_tests_tail.txt
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py > $O/r02_bench_cfg2_train.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; tail -2 $O/bench_cfg2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_cfg2_reference_arm.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --workload cfg2_eval --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_cfg2_eval.json 2> $O/bench_cfg2_eval.err; echo "cfg2_eval rc=$?"
python - <<'PY'
import json, glob
for f in ('gpurun_out/r02_bench_cfg2_train.json', 'gpurun_out/r02_bench_cfg2_reference_arm.json', 'gpurun_out/r02_bench_cfg2_eval.json'):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith('{')][-1]
        fw = d.get('fwd') or {}
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'Mpts/s', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'roof', (d.get('roofline') or {}).get('frac'),
              (d.get('roofline') or {}).get('traffic'), 'fwd', round(fw.get('value',0)/1e6,2), 'clk', d.get('clocks'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fwd"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 140 --csv --log-file $O/r02_launches_cfg2_train.csv $CMD > $O/ncu_l.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py $O/r02_launches_cfg2_train.csv > $O/r02_launch_summary_cfg2_train.txt 2>&1; head -12 $O/r02_launch_summary_cfg2_train.txt
CMDE="python bench.py --workload cfg2_eval --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $O/r02_launches_cfg2_eval.csv $CMDE > $O/ncu_le.log 2>&1
python tools/launch_summary.py $O/r02_launches_cfg2_eval.csv > $O/r02_launch_summary_cfg2_eval.txt 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 81 -c 27 -f -o /tmp/prof_gemm $CMD > $O/ncu_gemm.log 2>&1; echo "full gemm rc=$?"
python tools/ncu_summary.py /tmp/prof_gemm.ncu-rep > $O/r02_ncu_full_gemm_cfg2_train.txt 2>&1
timeout 900 ncu --set full --clock-control none -k regex:gemm_kernel -s 15 -c 5 -f -o /tmp/prof_gemm_eval $CMDE > $O/ncu_gemm_e.log 2>&1
python tools/ncu_summary.py /tmp/prof_gemm_eval.ncu-rep > $O/r02_ncu_full_gemm_cfg2_eval.txt 2>&1
python tools/ncu_traffic_lines.py $O/r02_ncu_full_gemm_cfg2_train.txt $O/r02_ncu_full_gemm_cfg2_eval.txt >> $O/r02_ncu_full_gemm_cfg2_train.txt
tail -8 $O/r02_ncu_full_gemm_cfg2_train.txt

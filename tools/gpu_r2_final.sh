#!/bin/bash
# round 2 final measurements on one B200: test-suite, smoke, bench lines, reference arm, ncu launch lists + full captures
mkdir -p gpurun_out
O=gpurun_out
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8 | cut -c1-300 | tee $O/r02_gpu_tests_tail.txt
echo "=== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py > $O/r02_bench_cfg2_train.json 2> $O/bench_cfg2.err; echo "cfg2 rc=$?"; tail -2 $O/bench_cfg2.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_cfg2_reference_arm.json 2> $O/bench_ref.err; echo "ref rc=$?"
for w in cfg2_eval cfg1_eval cfg3_train cfg3_eval cfg5_eval; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_$w.json 2> $O/bench_$w.err; echo "$w rc=$?"
done
timeout 900 python bench.py --workload cfg2_eval --precision bf16x3 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_cfg2_eval_bf16x3.json 2> $O/bench_x3.err; echo "x3 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench_*.json')):
    try:
        d = [json.loads(l) for l in open(f) if l.startswith('{')][-1]
        fw = d.get('fwd') or {}
        print(f.split('/')[-1], 'ms', round(d['ms_per_step'],4), 'Mpts/s', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), 'dtype', d.get('dtype'),
              'roof', (d.get('roofline') or {}).get('frac'), 'fwd', round(fw.get('value',0)/1e6,2), 'clk', d.get('clocks'), 'cpu', (d.get('cpu_baseline') or {}).get('value'),
              'eager', {k: round(v/1e6,2) for k, v in (d.get('torch_eager_same_gpu') or {}).items() if isinstance(v, float)}, d['config'].get('workload','')[:60])
    except Exception as e:
        print(f, 'ERR', e)
PY
# launch lists (one step each)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fwd"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 140 --csv --log-file $O/r02_launches_cfg2_train.csv $CMD > $O/ncu_l.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py $O/r02_launches_cfg2_train.csv > $O/r02_launch_summary_cfg2_train.txt 2>&1; head -40 $O/r02_launch_summary_cfg2_train.txt
CMDE="python bench.py --workload cfg2_eval --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 40 --csv --log-file $O/r02_launches_cfg2_eval.csv $CMDE > $O/ncu_le.log 2>&1
python tools/launch_summary.py $O/r02_launches_cfg2_eval.csv > $O/r02_launch_summary_cfg2_eval.txt 2>&1; head -14 $O/r02_launch_summary_cfg2_eval.txt
# full captures
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 78 -c 26 -f -o /tmp/prof_gemm $CMD > $O/ncu_gemm.log 2>&1; echo "full gemm rc=$?"
python tools/ncu_summary.py /tmp/prof_gemm.ncu-rep > $O/r02_ncu_full_gemm_cfg2_train.txt 2>&1
timeout 900 ncu --set full --clock-control none -k regex:gemm_kernel -s 15 -c 5 -f -o /tmp/prof_gemm_eval $CMDE > $O/ncu_gemm_e.log 2>&1
python tools/ncu_summary.py /tmp/prof_gemm_eval.ncu-rep > $O/r02_ncu_full_gemm_cfg2_eval.txt 2>&1
python tools/ncu_traffic_lines.py $O/r02_ncu_full_gemm_cfg2_train.txt $O/r02_ncu_full_gemm_cfg2_eval.txt >> $O/r02_ncu_full_gemm_cfg2_train.txt
cat $O/r02_ncu_full_gemm_cfg2_train.txt; cat $O/r02_ncu_full_gemm_cfg2_eval.txt
timeout 1200 ncu --set full --clock-control none -k regex:"k_bn_bwd_apply|k_bn_relu|k_head|k_ingest|k_convert|k_fill|k_adam|k_cloud|k_maxpool|k_predict|k_fold|k_gram|k_pool" -s 90 -c 40 -f -o /tmp/prof_ew $CMD > $O/ncu_ew.log 2>&1
python tools/ncu_summary.py /tmp/prof_ew.ncu-rep > $O/r02_ncu_full_pointwise_cfg2_train.txt 2>&1; cat $O/r02_ncu_full_pointwise_cfg2_train.txt

set -x
timeout 600 python -m pytest tests/test_gpu_tensordot.py tests/test_gpu_tcgen05.py tests/test_gpu_drivers.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python bench.py --networks 74 --steps 10 --no-cpu-baseline > gpurun_out/bench_n74_v8.json 2> gpurun_out/bench_n74_v8.err; tail -3 gpurun_out/bench_n74_v8.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n74_v8.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['result_check']); r=d['roofline']; print(r['bound'], r['achieved'], r['frac'], r['kernel'], r['kernel_share_of_step_time']); [print(k,v) for k,v in r['families'].items()]"
for c in cfg1 flagship cfg4 tree32; do timeout 300 python bench.py --config $c --steps 20 > gpurun_out/cfg_${c}_bf16.json 2> gpurun_out/cfg_${c}.err || tail -5 gpurun_out/cfg_${c}.err; done
timeout 300 python bench.py --config flagship --dtype f32 --steps 20 > gpurun_out/cfg_flagship_f32.json 2>> gpurun_out/cfg_flagship.err
timeout 300 python bench.py --config flagship --dtype f64 --steps 20 > gpurun_out/cfg_flagship_f64.json 2>> gpurun_out/cfg_flagship.err
timeout 300 python bench.py --config tree32 --dtype f32 --steps 20 > gpurun_out/cfg_tree32_f32.json 2>> gpurun_out/cfg_tree32.err
timeout 400 python bench.py --config cfg3 --dtype f64 --steps 2 > gpurun_out/cfg_cfg3_f64.json 2> gpurun_out/cfg_cfg3.err || tail -5 gpurun_out/cfg_cfg3.err
timeout 600 python bench.py --config cfg5 --dtype f64 --steps 6 > gpurun_out/cfg_cfg5_f64.json 2> gpurun_out/cfg_cfg5.err || tail -5 gpurun_out/cfg_cfg5.err
for f in gpurun_out/cfg_*.json; do echo $f; python -c "
import json,sys; d=json.load(open('$f')); print({k:d.get(k) for k in ('metric','value','ms_per_step','dtype')}); print(' roof', {k:(d.get('roofline') or {}).get(k) for k in ('bound','achieved','frac','kernel')}); print(' cpu', d.get('cpu_baseline')); print(' extra', {k:d.get(k) for k in ('parity','parity_ok','rel_err','rel_err_vs_fp64','sizes','site_update_seconds','energy_after_last_update')})"; done

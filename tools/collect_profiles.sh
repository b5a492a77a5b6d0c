#!/usr/bin/env bash
# copies the outputs of tools/gpu_profiles.sh from gpurun_out/ into profiles/ (text summaries only)
set -e
cd "$(dirname "$0")/.."
cp gpurun_out/bench_B.json profiles/r1_final_bench_B.json
cp gpurun_out/bench_B.err profiles/r1_final_layer_times_B.txt
cp gpurun_out/bench_ref.json profiles/r1_final_bench_reference_arm.json
cp gpurun_out/launches_net.csv profiles/r1_final_launches_net_B.csv
cp gpurun_out/pytest_gpu.log profiles/r1_final_pytest_gpu.log
cp gpurun_out/sam_sweep_final.log profiles/r1_final_softargmax_sweep.txt
grep -B1 'sam prof' gpurun_out/sam_prof_final.log > profiles/r1_final_softargmax_phases.txt || true
for c in C D E; do cp gpurun_out/bench_$c.json profiles/r1_final_bench_$c.json; done
for n in root b1u2 b3u2 b4u2 sam; do
  python tools/ncu_read.py gpurun_out/full_$n.ncu-rep 2>/dev/null | grep -E "^==|time_duration|dram__bytes_(read|write).sum |pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|lts__throughput.avg.pct|lts__t_sector_hit|gpu__dram_throughput.avg|launch__registers|launch__shared_mem_per_block_dynamic|launch__grid_size|launch__block_size|smsp__inst_executed.sum |issue_active.avg.pct|sm__warps_active.avg.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum" > profiles/r1_final_ncu_full_$n.txt
  python tools/ncu_stalls.py gpurun_out/full_$n.ncu-rep 12 2>/dev/null | head -80 > profiles/r1_final_ncu_stalls_$n.txt
done
python - <<'PY'
import csv, json
rows = list(csv.reader(open('gpurun_out/launches_net.csv')))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
kn, mn, mv, idc = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID'))
per = {}
for r in body:
    per.setdefault(int(r[idc]), {'name': r[kn]})[r[mn]] = float(r[mv].replace(',', ''))
step = sorted(per)[-52:]
b = lambda i: per[i]['dram__bytes_read.sum'] + per[i]['dram__bytes_write.sum']
conv_b = sum(b(i) for i in step if 'conv_gemm' in per[i]['name'])
tot_b = sum(b(i) for i in step)
tot_t = sum(per[i]['gpu__time_duration.sum'] for i in step)
print('step: kernel time %.1f us (ncu, serialised), dram %.2f GB (conv_gemm %.2f GB)' % (tot_t / 1e3, tot_b / 1e9, conv_b / 1e9))
sam_b = sum(b(i) for i in step if 'softargmax' in per[i]['name'])
json.dump({'workload': 'config B, 256 crops', 'conv_gemm_launches': 49, 'conv_gemm_dram_bytes_per_step': conv_b, 'step_dram_bytes': tot_b,
           'softargmax_dram_bytes_per_launch': sam_b,
           'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, second step of tools/ncu_targets.py net (profiles/r1_final_launches_net_B.csv)'},
          open('profiles/traffic_B.json', 'w'), indent=1)
for c in 'BDCE':
    d = json.loads(open(f'gpurun_out/bench_{c}.json').read().strip().splitlines()[-1])
    print(d['config']['workload'], '| value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'conv frac',
          round(d['roofline']['frac'], 3), 'step frac', round(d['config']['tensor_frac_whole_step'], 3), 'sam GB/s', round(d['roofline_softargmax']['achieved']), d['clocks']['reasons'])
PY
tail -1 gpurun_out/pytest_gpu.log

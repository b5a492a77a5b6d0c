#!/usr/bin/env bash
# copies the outputs of tools/gpu_r2_profiles.sh from gpurun_out/ into profiles/ (text summaries only); `TAG` names the set
set -e
cd "$(dirname "$0")/.."
TAG=${TAG:-r2_final}
cp gpurun_out/bench_B.json profiles/${TAG}_bench_B.json
cp gpurun_out/bench_B.err profiles/${TAG}_layer_times_B.txt
[ -f gpurun_out/bench_ref.json ] && cp gpurun_out/bench_ref.json profiles/${TAG}_bench_reference_arm.json
cp gpurun_out/pytest_gpu.log profiles/${TAG}_pytest_gpu.log
cp gpurun_out/parity_table.json profiles/${TAG}_parity_table.json
cp gpurun_out/sam_sweep_final.log profiles/${TAG}_softargmax_sweep.txt
grep -B1 'sam prof' gpurun_out/sam_prof_final.log > profiles/${TAG}_softargmax_phases.txt || true
cp gpurun_out/roles.txt profiles/${TAG}_role_timers_B.txt
for c in A C D E; do [ -f gpurun_out/bench_$c.json ] && cp gpurun_out/bench_$c.json profiles/${TAG}_bench_$c.json; [ -f gpurun_out/bench_$c.err ] && grep " us$" gpurun_out/bench_$c.err > profiles/${TAG}_layer_times_$c.txt; done
for c in B C D E; do [ -f gpurun_out/launches_net_$c.csv ] && cp gpurun_out/launches_net_$c.csv profiles/${TAG}_launches_net_$c.csv; done
for n in root chain chain3 b3c2 sam; do
  [ -f gpurun_out/full_$n.ncu-rep ] || continue
  python tools/ncu_read.py gpurun_out/full_$n.ncu-rep 2>/dev/null | grep -E "^==|time_duration|dram__bytes_(read|write).sum |pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|lts__throughput.avg.pct|lts__t_sector_hit|gpu__dram_throughput.avg|launch__registers|launch__shared_mem_per_block_dynamic|launch__grid_size|launch__block_size|smsp__inst_executed.sum |issue_active.avg.pct|sm__warps_active.avg.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum" > profiles/${TAG}_ncu_full_$n.txt
  python tools/ncu_stalls.py gpurun_out/full_$n.ncu-rep 12 2>/dev/null | head -80 > profiles/${TAG}_ncu_stalls_$n.txt
done
for t in memcheck racecheck synccheck; do for w in net ops strict; do [ -f gpurun_out/sanitizer_${t}_${w}.log ] && echo "== $t $w: $(grep -h 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/sanitizer_${t}_${w}.log | tail -1)"; done; done > profiles/${TAG}_sanitizer_lines.txt
python - <<PY
import csv, json, os
tag = os.environ.get('TAG', 'r2_final')
for c in 'BCDE':
    path = f'gpurun_out/launches_net_{c}.csv'
    if not os.path.exists(path):
        continue
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]; body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    kn, mn, mv, idc = (hdr.index(k) for k in ('Kernel Name', 'Metric Name', 'Metric Value', 'ID'))
    per = {}
    for r in body:
        per.setdefault(int(r[idc]), {'name': r[kn]})[r[mn]] = float(r[mv].replace(',', ''))
    ids = sorted(per)
    step = ids[len(ids) // 2:]                      # the second of the two steps
    b = lambda i: per[i]['dram__bytes_read.sum'] + per[i]['dram__bytes_write.sum']
    conv = [i for i in step if 'conv_gemm' in per[i]['name'] or 'conv_chain' in per[i]['name']]
    conv_b = sum(b(i) for i in conv); tot_b = sum(b(i) for i in step)
    conv_t = sum(per[i]['gpu__time_duration.sum'] for i in conv); tot_t = sum(per[i]['gpu__time_duration.sum'] for i in step)
    sam_b = sum(b(i) for i in step if 'softargmax' in per[i]['name'])
    print(f'config {c}: {len(step)} launches/step, kernel time {tot_t / 1e3:.1f} us (ncu, serialised; convolutions {conv_t / 1e3:.1f} us = {conv_t / tot_t:.3f} of it), dram {tot_b / 1e9:.2f} GB (convolutions {conv_b / 1e9:.2f} GB)')
    json.dump({'workload': f'config {c}', 'conv_launches': len(conv), 'conv_gemm_dram_bytes_per_step': conv_b, 'step_dram_bytes': tot_b,
               'softargmax_dram_bytes_per_launch': sam_b, 'conv_ns_ncu': conv_t, 'step_ns_ncu': tot_t,
               'source': f'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum, second step of tools/ncu_targets.py net_cfg {c} (profiles/{tag}_launches_net_{c}.csv)'},
              open(f'profiles/traffic_{c}.json', 'w'), indent=1)
for c in 'BDCEA':
    p = f'gpurun_out/bench_{c}.json'
    if not os.path.exists(p):
        continue
    d = json.loads(open(p).read().strip().splitlines()[-1])
    print(d['config']['workload'], '| value', round(d['value']), 'ms', round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'conv frac',
          round(d['roofline']['frac'], 3), 'step frac', round(d['details']['tensor_frac_whole_step'], 3), 'sam GB/s', round(d['roofline_softargmax']['achieved']), d['clocks']['reasons'])
PY
tail -1 gpurun_out/pytest_gpu.log

"""turn the .ncu-rep files of gpurun_out/ into the text summaries committed under profiles/ (+ ncu_traffic.json)
    python tools/make_profiles.py r01_v7"""
import csv, json, os, re, subprocess, sys
tag = sys.argv[1]
os.makedirs('profiles', exist_ok=True)
PAT = re.compile(r'gpu__time_duration.sum|registers_per_thread$|warps_active.avg.pct|sm__throughput.avg.pct|inst_executed.sum$|'
                 r'dram__bytes_(read|write).sum$|shared_mem_per_block$|sm__cycles_elapsed.max$|sm__pipe_tensor.*cycles_active.avg.pct_of_peak_sustained_elapsed|'
                 r'gpu__dram_throughput.avg.pct|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|lts__t_sector_hit_rate.pct|waves_per|'
                 r'sm__inst_executed_pipe_(fma|lsu|uniform|tma|tc).*sum$|smsp__inst_executed_op_(ldgsts|shared).*sum$')
traffic = {}
for f in sorted(os.listdir('gpurun_out')):
    if not (f.startswith('prof_') and f.endswith('.ncu-rep')):
        continue
    name = f[5:-8]
    rep = os.path.join('gpurun_out', f)
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    lines = [f'# ncu --set full --clock-control none (one launch, cold cache, serialised) : {name}']
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"==== {d['Kernel Name'][:90]}  grid {d['Grid Size']} block {d['Block Size']}")
        rd = wr = 0.0
        for h, u, v in zip(hdr, units, r):
            if PAT.search(h) and v not in ('', '0', '0.00'):
                lines.append(f'  {h} = {v} {u}')
            if h == 'dram__bytes_read.sum' or h == 'dram__bytes_write.sum':
                x = float(v.replace(',', '')) if v else 0.0
                x *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
                if h.endswith('read.sum'): rd = x
                else: wr = x
        m = re.match(r'(.*)_B(\d+)$', name)
        if m:
            key = {'node_fwd': 'bmnas_node_fwd', 'node_bwd': 'bmnas_node_bwd', 'sg_fwd': 'bmnas_conv_fwd', 'sg_dgrad': 'bmnas_conv_dgrad',
                   'wgrad': 'bmnas_conv_wgrad', 'mix_bwd': 'bmnas_mix_bwd', 'ln_bwd': 'bmnas_ln_bwd', 'panel_fwd': 'bmnas_conv_fwd',
                   'node_fwd_warp': 'bmnas_node_fwd', 'node_bwd_warp': 'bmnas_node_bwd', 'mixed_fwd': 'bmnas_mixed_fwd', 'mixed_fwd_tf32': 'bmnas_mixed_fwd',
                   'mixed_fwd_bf16': 'bmnas_mixed_fwd_bf16', 'panel_dgrad': 'bmnas_conv_dgrad', 'tc_wgrad': 'bmnas_conv_wgrad', 'ws_dgrad': 'bmnas_conv_dgrad', 'ws_fwd': 'bmnas_conv_fwd',
                   'ws_wgrad': 'bmnas_conv_wgrad', 'mixed_small': 'bmnas_mixed_small_fwd', 'head': 'bmnas_head_fused'}.get(m.group(1), m.group(1))
            traffic[f'{key}@B{m.group(2)}'] = int(rd + wr)
    src = subprocess.run([sys.executable, 'tools/ncu_src.py', rep, '14'], capture_output=True, text=True).stdout
    lines.append('# hottest SASS lines (sampled stalls)')
    lines += src.splitlines()
    open(f'profiles/{tag}_ncu_{name}.txt', 'w').write('\n'.join(lines) + '\n')
    print('wrote', f'profiles/{tag}_ncu_{name}.txt')
if traffic:
    p = 'profiles/ncu_traffic.json'
    old = json.load(open(p)) if os.path.exists(p) else {}
    old.update(traffic)
    # which build the captures belong to (bench.py prints it next to roofline.traffic)
    old['_git'] = subprocess.run(['git', 'rev-parse', '--short', 'HEAD'], capture_output=True, text=True).stdout.strip() + \
        ('+dirty' if subprocess.run(['git', 'status', '--porcelain', '--', 'bm-nas_b200/csrc', 'include'], capture_output=True, text=True).stdout.strip() else '')
    old['_tag'] = tag
    json.dump(old, open(p, 'w'), indent=1, sort_keys=True)
    print(traffic)
for src, dst in (('bench.log', f'{tag}_bench.json'), ('bench_reference.log', f'{tag}_bench_reference.json'), ('launch_list.txt', f'{tag}_launch_list.txt')):
    if os.path.exists('gpurun_out/' + src):
        open('profiles/' + dst, 'w').write(open('gpurun_out/' + src).read())
if os.path.exists('gpurun_out/bench_full.log'):
    open(f'profiles/{tag}_kernel_times.txt', 'w').write(''.join(l for l in open('gpurun_out/bench_full.log') if not l.startswith('{')))

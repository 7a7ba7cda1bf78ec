"""Pretty-print gpurun_out/probe9.log-style timing lines and trace summaries."""
import json, sys
for path in sys.argv[1:]:
    for l in open(path):
        if not l.startswith('{'):
            continue
        d = json.loads(l)
        if 'gflops' in d:
            print(d['spec'].get('tag'), d['spec']['M'], {k: (round(v / 1e3, 1) if isinstance(v, float) else v) for k, v in d['gflops'].items()})
        elif 'mainloop' in d:
            print(d['id'], d['N'], d['tag'], 'span', d['span_us'], 'hdr', d['hdr'])
            for k in ('mainloop', 'epilogue'):
                print('  ', k, {c: (v['n'], v['mean_us'], v['min_us'], v['max_us']) for c, v in d[k].items() if v})
            print('   check', d['check_phase'])
            print('   gap', d['mma_gap_between_items'], 'acclag', d['acc_done_after_last_issue'])
            print('   unit_end', d['unit_end_us'], 'first_epi', d['first_data_epilogue_start_us'], 'enc_end', d.get('encode_end_us'), 'enc_dur', d.get('encode_dur_us'), 'enc_total', d.get('encode_all_done_after_first_start_us'), 'worker0(wait,work,slots)', d.get('enc_worker0_wait_work_us_slots'), 'chk_end', d['chk_items_end_us'][:6])
        elif 'stats' in d:
            print('stats', {k: d['stats'][k] for k in ('detected', 'max_rel_residual')})

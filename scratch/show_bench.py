"""print the interesting numbers of one or more bench.py JSON lines (scratch helper)"""
import json, sys
def load(p):
    for ln in reversed(open(p).read().splitlines()):
        if ln.startswith('{'):
            return json.loads(ln)
for p in sys.argv[1:]:
    l = load(p)
    if not l:
        print(p, 'unparsed'); continue
    print('==', p, 'N =', l['n_gpus'])
    if 'parity_check' in l:
        print('  parity', {k: v for k, v in l['parity_check'].items() if k in ('ok', 'max_rel', 'ranks', 'failed')})
    print('  value %.1f M/s  ms %.2f  roofline.frac %.3f dram_frac %s' % (l['value'] / 1e6, l['ms_per_step'], l['roofline']['frac'], l['roofline'].get('dram_frac')))
    if 'e2e' in l:
        e = l['e2e']
        print('  e2e pinned %.1f  pageable %.1f  registered %.1f (reg %.2fs)' % (e['value'] / 1e6, e['pageable']['value'] / 1e6, e['registered']['value'] / 1e6, e['registered']['register_seconds_once']))
        print('  e2e_predict %.1f pageable %.1f' % (l['e2e_predict']['value'] / 1e6, l['e2e_predict']['pageable']['value'] / 1e6))
    if 'c4_decision_function' in l:
        print('  c4 predict %.1f' % (l['c4_decision_function']['samples_per_s'] / 1e6))
    if 'c4_adagrad' in l:
        print('  c4 adagrad %.1f (frac %.2f)' % (l['c4_adagrad']['samples_per_s'] / 1e6, l['c4_adagrad']['frac']))
    if 'c3_mbpsgd' in l:
        for k in ('reference_default', '256Ki_per_gpu', '1Mi_per_gpu'):
            v = l['c3_mbpsgd'][k]
            print('  c3 %-18s %.1f M/s  %.0f us/mb  frac %.2f' % (k, v['samples_per_s'] / 1e6, v['us_per_minibatch'], v['frac_sparse']))
    if 'c5_ffm' in l:
        print('  c5', {k: round(v['samples_per_s'] / 1e6, 2) for k, v in l['c5_ffm'].items() if isinstance(v, dict)})
    if 'cd' in l:
        print('  cd c1 %.2f ms c2 %.2f ms' % (l['cd']['c1']['epoch_s'] * 1e3, l['cd']['c2']['epoch_s'] * 1e3))
    if 'uniform' in l:
        print('  uniform grad %.1f (frac %.3f) fwd %.1f' % (l['uniform']['grad']['samples_per_s'] / 1e6, l['uniform']['grad']['frac'], l['uniform']['forward']['samples_per_s'] / 1e6))

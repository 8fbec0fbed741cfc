"""Pretty-print the JSON line(s) of bench.py: python tools/show_bench.py <file> (or stdin)."""
import signal
import sys, json

signal.signal(signal.SIGPIPE, signal.SIG_DFL)  # `| head` is fine
for line in open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin:
    line = line.rstrip()
    if line.startswith('{"metric"') or line.startswith('{"impl"'):
        d = json.loads(line)
        print('BENCH', d['config']['workload'][:70], 'value=%.3e ms/step=%.2f launches=%s' % (d['value'], d['ms_per_step'], d.get('gpu_launches')))
        for k, v in d.get('stages', {}).items():
            print('   %-12s %8.3f ms share %.3f %s' % (k, v['ms_per_call'], v['share'], ('frac=%.3f' % v['frac']) if 'frac' in v else ''))
        for k in ('gemm', 'roofline', 'e2e', 'cpu_baseline', 'clocks'):
            if k in d: print('   ', k, d[k])
    else:
        print(line[:220])

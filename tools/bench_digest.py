"""One-screen digest of a bench.py JSON line."""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split('\n')[-1])
print('n_gpus %d | value %.4e terms/s | %.3f ms/step | e2e %.4e (%.3f ms/step) | launches %d | clocks %s' % (
    d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches'], d['clocks']))
print('checks', d.get('checks'))
print('steps_per_s', d.get('steps_per_s'))
r = d['roofline']
print('roofline frac %.3f (nominal) | fp64_pipe_frac %s | kernel_ms %.4f | traffic %s | grid %s' % (
    r['frac'], r['fp64_pipe_frac'], r['kernel_ms'], r['traffic'], r['grid']))

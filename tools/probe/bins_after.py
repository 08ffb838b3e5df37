import sys, os, time, gc
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import config_sweep as cs
name, model, truth, nw = cs.config_c2()
print(cs.run(name, model, truth, nw, True)[:80])
model.pack().close(); del model
for i in range(4):
    t0 = time.perf_counter(); line = cs.bins_workflow(); print('%.3f s total |' % (time.perf_counter() - t0), line[60:140])
    if i == 1:
        gc.collect()

import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import config_sweep as cs
for _ in range(3):
    print(cs.bins_workflow())

# Timeline of one likelihood launch on the small BASELINE configurations; needs a library built with
# -DMCD_KERNEL_PROFILE (see csrc/mcd_kernels.cu) in MCD_B200_LIB.
for c in c1 c2 c4; do echo "== $c"; timeout 100 python -c "
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tools'))
import config_sweep as cs
from mcmc_dynamics_b200 import synthetic
name, model, truth, nw = {'c1': cs.config_c1, 'c2': cs.config_c2, 'c4': cs.config_c4}['$c']()
th = synthetic.initial_ball(truth, model.fitted_parameters, nw, seed=5, scale=0.05)[:nw // 2]
for _ in range(6): model.lnprob(th)
" 2>&1 | grep "^launch" | tail -2; done

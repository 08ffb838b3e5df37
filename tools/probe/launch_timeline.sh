# Timeline of one likelihood launch on the BASELINE configurations; needs a library built with
# -DMCD_KERNEL_PROFILE (python tools/build_variant.py profile -DMCD_KERNEL_PROFILE) in MCD_B200_LIB.
export MCD_B200_LIB=${MCD_B200_LIB:-scratch_ab/profile/libmcd_b200.so}
for c in ${@:-c1 c2 c3 c4}; do echo "== $c"; timeout 200 python -c "
import sys, os
sys.path.insert(0, os.getcwd())
from mcmc_dynamics_b200 import configs, synthetic
name, model, truth, nw = configs.BUILDERS['$c'.upper().replace('C3B', 'C3b')]()
th = synthetic.initial_ball(truth, model.fitted_parameters, nw, seed=5, scale=0.05)[:nw // 2]
for _ in range(6): model.lnprob(th)
" 2>&1 | grep "^launch" | tail -2; done

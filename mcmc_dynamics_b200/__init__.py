"""B200-native likelihood path of mcmc-dynamics: the reference's model-class interface
(``ConstantFit``, ``ModelFit`` and their background variants, ``Parameters``, ``DataReader``) on top
of hand-written sm_100a CUDA kernels reached through a C ABI (``include/mcd_b200.h``)."""
from .parameter import Parameter, Parameters
from .data_reader import DataReader

__all__ = ['Parameter', 'Parameters', 'DataReader']
__version__ = '0.1.0'

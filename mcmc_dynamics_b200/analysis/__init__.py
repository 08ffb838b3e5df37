from .runner import Runner
from .constant import ConstantFit, ConstantFitGB
from .model import ModelFit, ModelFitGB, ModelFitConstantBackground
from .bins import RadialBinsFit

__all__ = ['Runner', 'ConstantFit', 'ConstantFitGB', 'ModelFit', 'ModelFitGB', 'ModelFitConstantBackground',
           'RadialBinsFit']

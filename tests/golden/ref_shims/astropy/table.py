"""Stand-in for astropy.table.QTable / Table: named equal-length columns."""
from collections import OrderedDict

import numpy as np


class Table(object):
    def __init__(self, data=None, names=None, **kwargs):
        self.columns = OrderedDict()
        if data is None:
            return
        if isinstance(data, Table):
            for name, col in data.columns.items():
                self.columns[name] = col
        elif isinstance(data, dict):
            for name, col in data.items():
                self.columns[name] = col if hasattr(col, 'unit') else np.asarray(col)
        else:
            for name, col in zip(names, data):
                self.columns[name] = col if hasattr(col, 'unit') else np.asarray(col)

    @property
    def colnames(self):
        return list(self.columns)

    def __len__(self):
        for col in self.columns.values():
            return len(col)
        return 0

    def __contains__(self, name):
        return name in self.columns

    def __getitem__(self, item):
        if isinstance(item, str):
            return self.columns[item]
        out = type(self)()
        for name, col in self.columns.items():
            out.columns[name] = col[item]
        return out

    def __setitem__(self, name, values):
        self.columns[name] = values if hasattr(values, 'unit') else np.asarray(values)

    def add_row(self, *args, **kwargs):
        raise NotImplementedError


class QTable(Table):
    pass

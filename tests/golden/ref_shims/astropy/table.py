"""Stand-in for astropy.table.QTable / Table: named equal-length columns, plus the little bit of
indexing (``add_index``, ``loc``) and column building the reference's post-processing uses."""
from collections import OrderedDict

import numpy as np


class _Column(object):
    """``QTable.Column(values, name=..., unit=...)``: values plus a name."""

    def __new__(cls, values, name=None, unit=None):
        if unit is not None and not hasattr(values, 'unit'):
            from astropy import units as u
            values = u.Quantity(values, unit)
        elif not hasattr(values, 'unit'):
            values = np.asarray(values)
        holder = object.__new__(cls)
        holder.values = values
        holder.name = name
        return holder


class _Row(object):
    def __init__(self, table, index):
        self._table = table
        self._index = index

    def __getitem__(self, name):
        col = self._table.columns[name]
        if hasattr(col, 'unit'):
            from astropy import units as u
            return u.Quantity(col.value[self._index], col.unit)
        return col[self._index]

    def __setitem__(self, name, value):
        col = self._table.columns[name]
        if hasattr(col, 'unit') and hasattr(value, 'unit'):
            value = value.to(col.unit).value
        col[self._index] = value


class _Loc(object):
    def __init__(self, table):
        self._table = table

    def __getitem__(self, key):
        keys = list(self._table.columns[self._table._index_column])
        return _Row(self._table, keys.index(key))


class Table(object):
    Column = _Column

    def __init__(self, data=None, names=None, **kwargs):
        self.columns = OrderedDict()
        self._index_column = None
        if data is None:
            return
        if isinstance(data, Table):
            for name, col in data.columns.items():
                self.columns[name] = col
        elif isinstance(data, dict):
            for name, col in data.items():
                self.columns[name] = col if hasattr(col, 'unit') else np.asarray(col)
        elif names is None:                       # Table([Column(...), Column(...)])
            for col in data:
                self.columns[col.name] = col.values
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

    def add_index(self, name):
        self._index_column = name

    @property
    def loc(self):
        return _Loc(self)

    def add_column(self, col, name=None):
        if isinstance(col, _Column):
            self.columns[name or col.name] = col.values
        else:
            self.columns[name] = col if hasattr(col, 'unit') else np.asarray(col)


class QTable(Table):
    pass

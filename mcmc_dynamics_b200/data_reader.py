"""``DataReader``: the star-catalogue container the model classes consume.

Mirrors ``mcmc_dynamics/utils/files/data_reader.py:10-140`` (a thin wrapper around an
``astropy.table.QTable``) on top of a minimal column table, because astropy is
not a dependency here.  Columns are float64 arrays with an optional unit; the
model classes read ``v, verr`` [km/s], ``ra, dec`` [deg] and optionally
``pmember``, ``density`` (``analysis/runner.py:75-81,103``).
"""
import logging
from collections import OrderedDict

import numpy as np

from . import units as u

logger = logging.getLogger(__name__)


class Column(u.Quantity):
    """A table column: values plus unit.  Unit-less columns report the dimensionless unit."""

    def __getitem__(self, item):
        out = np.asarray(self.value)[item]
        if np.ndim(out) == 0:
            return u.Quantity(out, self.unit) if not self.unit.is_unity() else out
        return Column(out, self.unit)

    def min(self):
        return np.min(self.value) if self.unit.is_unity() else u.Quantity(np.min(self.value), self.unit)

    def max(self):
        return np.max(self.value) if self.unit.is_unity() else u.Quantity(np.max(self.value), self.unit)


class Table(object):
    """Just enough of ``QTable``: named equal-length columns, row masks, ``len`` and ``columns``."""

    def __init__(self, data=None, names=None, units=None):
        self._columns = OrderedDict()
        if data is None:
            return
        if isinstance(data, Table):
            for name in data.columns:
                self[name] = data[name]
        elif hasattr(data, 'colnames'):                 # astropy Table / QTable
            for name in data.colnames:
                self[name] = data[name]
        elif hasattr(data, 'columns') and hasattr(data, 'to_numpy'):     # pandas DataFrame
            for name in data.columns:
                self[str(name)] = data[name].to_numpy()
        elif isinstance(data, dict):
            for name, values in data.items():
                self[name] = values
        elif isinstance(data, np.ndarray) and data.dtype.names:
            for name in data.dtype.names:
                self[name] = data[name]
        else:
            columns = list(data)
            if names is None:
                raise ValueError('Column names are required to build a table from a list of columns.')
            for name, values in zip(names, columns):
                self[name] = values
        if units:
            for name, unit in units.items():
                self[name] = u.Quantity(self[name].value, unit)

    @property
    def columns(self):
        return self._columns

    @property
    def colnames(self):
        return list(self._columns)

    def __len__(self):
        for col in self._columns.values():
            return len(col)
        return 0

    def __contains__(self, name):
        return name in self._columns

    def __setitem__(self, name, values):
        if u.is_quantity(values):
            q = u.as_quantity(values)
            col = Column(np.atleast_1d(np.array(q.value, dtype=np.float64)), q.unit)
        else:
            arr = np.atleast_1d(np.asarray(values))
            if arr.dtype.kind in 'iub':
                col = _PlainColumn(arr)
            else:
                col = Column(np.array(arr, dtype=np.float64), u.dimensionless_unscaled)
        if self._columns and name not in self._columns and len(col) != len(self):
            raise ValueError("Inconsistent data column lengths: {0} vs {1}".format(len(col), len(self)))
        self._columns[name] = col

    def __getitem__(self, item):
        if isinstance(item, str):
            return self._columns[item]
        # row selection by boolean mask, index array or slice
        if u.is_quantity(item):
            item = np.asarray(item.value)
        out = Table()
        for name, col in self._columns.items():
            out._columns[name] = col[item] if not isinstance(col, _PlainColumn) else _PlainColumn(col.value[item])
        return out

    def __repr__(self):
        head = ' '.join('{0}[{1}]'.format(n, c.unit) for n, c in self._columns.items())
        return '<Table length={0}: {1}>'.format(len(self), head)


class _PlainColumn(object):
    """Integer / boolean column (e.g. the radial ``bin`` index): no unit, numpy semantics."""

    unit = u.dimensionless_unscaled

    def __init__(self, values):
        self.value = np.asarray(values)

    def __len__(self):
        return len(self.value)

    def __getitem__(self, item):
        return self.value[item]

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.value, dtype=dtype)

    def __eq__(self, other):
        return self.value == other

    def __ne__(self, other):
        return self.value != other

    __hash__ = None

    def min(self):
        return self.value.min()

    def max(self):
        return self.value.max()


def calc_xy_offset(ra, dec, ra_center, dec_center):
    """Tangent-plane offsets (dx, dy) in arcmin of (ra, dec) from a centre, all angles in degrees
    unless they carry a unit (``utils/coordinates/calc_xy_offset.py:9-33``, van de Ven+ 2006).

    Host-side helper used for binning and for the one-time pack of fixed-centre fits; the
    per-walker geometry of free-centre fits is evaluated inside the CUDA kernel.
    """
    ra = np.deg2rad(u.strip(ra, u.deg))
    dec = np.deg2rad(u.strip(dec, u.deg))
    ra_center = np.deg2rad(u.strip(ra_center, u.deg))
    dec_center = np.deg2rad(u.strip(dec_center, u.deg))
    r0 = 10800. / np.pi
    dx = -r0 * np.cos(dec) * np.sin(ra - ra_center)
    dy = r0 * (np.sin(dec) * np.cos(dec_center) - np.cos(dec) * np.sin(dec_center) * np.cos(ra - ra_center))
    return u.Quantity(dx, u.arcmin), u.Quantity(dy, u.arcmin)


class DataReader(object):

    def __init__(self, data, **kwargs):
        """
        Parameters
        ----------
        data : dict, Table, structured ndarray, pandas DataFrame or astropy table
            The star catalogue, one entry per column.
        kwargs
            Passed on to :class:`Table` (``names``, ``units``).
        """
        self.data = data if isinstance(data, Table) and not kwargs else Table(data, **kwargs)

    @property
    def sample_size(self):
        return len(self.data)

    @property
    def has_ra(self):
        return 'ra' in self.data.columns

    @property
    def has_dec(self):
        return 'dec' in self.data.columns

    @property
    def has_coordinates(self):
        return self.has_ra & self.has_dec

    def compute_distances(self, ra_center, dec_center):
        """Distances of the stars from a reference point (data_reader.py:46-69)."""
        if not self.has_coordinates:
            logger.error('Cannot calculate distances as world coordinates are missing.')
            return
        x, y = calc_xy_offset(self.data['ra'], self.data['dec'], ra_center, dec_center)
        return u.Quantity(np.sqrt(x.value ** 2 + y.value ** 2), u.arcmin)

    def make_radial_bins(self, ra_center, dec_center, nstars=50, dlogr=0.2):
        """Assign each star to a radial bin holding at least `nstars` stars and spanning at least
        `dlogr` in log10(radius); the result goes to the ``bin`` column (data_reader.py:71-120)."""
        if not self.has_coordinates:
            logger.error('Cannot create radial profile. WCS coordinates of data points unknown.')
            return

        r = self.compute_distances(ra_center, dec_center).value
        order = np.argsort(r)
        log_r = np.log10(r[order])
        n = self.sample_size

        labels = -np.ones(n, dtype=np.int16)
        current = -1
        start = 0
        while start < (n - nstars):
            stop = min(n, start + nstars)
            # grow the bin until it is wide enough in log-radius
            while (log_r[stop] - log_r[start]) < dlogr:
                stop += 1
                if stop >= n:
                    break
            current += 1
            labels[start:stop] = current
            start = stop

        # leftover stars form their own bin if there are enough of them, else join the last one
        if (n - start) > 0.5 * nstars or current == -1:
            labels[start:] = current + 1
        else:
            labels[start:] = current

        self.data['bin'] = labels[order.argsort()]

    def fetch_radial_bin(self, i):
        """A new reader holding only the stars of bin `i` (data_reader.py:122-140)."""
        if 'bin' not in self.data.columns:
            logger.error('No information about bins available.')
            return None
        elif i < self.data['bin'].min() or i > self.data['bin'].max():
            logger.error('Requested bin {0} does not exist.'.format(i))
            return None

        return self.__class__(self.data[self.data['bin'] == i])

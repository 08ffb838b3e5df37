#!/usr/bin/env python
"""Input fixture of BASELINE.json config 1: the reference's example catalogue
(/root/reference/example/data/test.csv: rows r [arcmin], theta [rad], v, verr [km/s]; 6284 stars) placed
on the sky around the centre used by bin/run_tests.py:44 so that the coordinate-based model classes can
consume it.  theta is the position angle in the tangent plane measured like the reference's
theta = atan2(dy, dx), so (dx, dy) = r (cos theta, sin theta); the inverse gnomonic-like projection of
calc_xy_offset.py:30-31 is solved exactly for (ra, dec).  Output: c1_example_catalogue.npz (inputs only).
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
RA0, DEC0 = 56.345, -26.675


def main():
    r, theta, v, verr = np.loadtxt('/root/reference/example/data/test.csv', delimiter=',')
    r0 = 10800. / np.pi
    dx, dy = r * np.cos(theta) / r0, r * np.sin(theta) / r0     # in units of r0
    # calc_xy_offset: dx = -cos(dec) sin(dra); dy = sin(dec) cos(dec0) - cos(dec) sin(dec0) cos(dra)
    # with w = cos(dec) cos(dra) = sqrt(1 - dx^2 - dy^2)-like third component of the rotated unit vector
    d0 = np.deg2rad(DEC0)
    w = np.sqrt(1.0 - dx ** 2 - dy ** 2)
    sin_dec = dy * np.cos(d0) + w * np.sin(d0)
    dec = np.arcsin(sin_dec)
    cos_dec_cos_dra = w * np.cos(d0) - dy * np.sin(d0)
    dra = np.arctan2(-dx, cos_dec_cos_dra)
    ra = RA0 + np.rad2deg(dra)
    np.savez_compressed(os.path.join(HERE, 'c1_example_catalogue.npz'), ra=ra, dec=np.rad2deg(dec), v=v, verr=verr,
                        r_arcmin=r, theta=theta, ra_center=RA0, dec_center=DEC0)
    # self-check against the projection it inverts
    dec_r, dra_r = dec, np.deg2rad(ra - RA0)
    bx = -r0 * np.cos(dec_r) * np.sin(dra_r)
    by = r0 * (np.sin(dec_r) * np.cos(d0) - np.cos(dec_r) * np.sin(d0) * np.cos(dra_r))
    print('max |dx - r cos(theta)|, |dy - r sin(theta)| [arcmin]:', np.abs(bx - r * np.cos(theta)).max(),
          np.abs(by - r * np.sin(theta)).max())


if __name__ == '__main__':
    main()

"""Writes tests/golden/bsc5_three_stars.{le,be}.bin: a BSC5-format binary catalog built BY HAND, byte by byte with
`struct.pack`, from the published record layout of the Yale Bright Star Catalog's binary distribution (the file
/root/reference/constants.py:33 `BSC5BIN` points at, absent from the mount) — independent of the numpy record dtype the
loader uses.

Header, 28 bytes: seven 4-byte integers STAR0 = 0, STAR1 = 1, STARN = -3 (negative: J2000 positions), STNUM = 1,
MPROP = 1, NMAG = 1, NBENT = 32.  Entry, 32 bytes: XNO real*4 (HR number), SRA0 real*8 [rad], SDEC0 real*8 [rad],
IS char*2 (spectral type), MAG int*2 (V x 100), XRPM real*4, XDPM real*4 [rad/yr].

Entries (J2000, the catalog's values to the precision quoted):
  HR 2491  Sirius   06h45m08.9s  -16d42'58"   V -1.46  A1   pm -0.546, -1.223 "/yr
  HR 7001  Vega     18h36m56.3s  +38d47'01"   V +0.03  A0   pm +0.201, +0.287 "/yr
  HR    1           00h05m09.9s  +45d13'45"   V +6.70  A1   pm -0.012, -0.018 "/yr
plus one position-less entry (RA = Dec = 0, as the catalog's 14 removed objects have) that the loader must drop.
"""
import math
import os
import struct

ARCSEC = math.pi / 180.0 / 3600.0


def hms(h, m, s):
    return (h + m / 60.0 + s / 3600.0) * 15.0 * math.pi / 180.0


def dms(sign, d, m, s):
    return sign * (d + m / 60.0 + s / 3600.0) * math.pi / 180.0


STARS = [  # (HR, ra [rad], dec [rad], spectral, V*100, pm_ra [rad/yr], pm_dec [rad/yr])
    (2491, hms(6, 45, 8.9), dms(-1, 16, 42, 58.0), b"A1", -146, -0.546 * ARCSEC, -1.223 * ARCSEC),
    (7001, hms(18, 36, 56.3), dms(+1, 38, 47, 1.0), b"A0", 3, 0.201 * ARCSEC, 0.287 * ARCSEC),
    (92, 0.0, 0.0, b"  ", 0, 0.0, 0.0),                                   # removed object: no position
    (1, hms(0, 5, 9.9), dms(+1, 45, 13, 45.0), b"A1", 670, -0.012 * ARCSEC, -0.018 * ARCSEC),
]


def build(order: str) -> bytes:
    out = struct.pack(order + "7i", 0, 1, -len(STARS), 1, 1, 1, 32)
    for hr, ra, dec, sp, mag, pra, pdec in STARS:
        out += struct.pack(order + "f", float(hr)) + struct.pack(order + "d", ra) + struct.pack(order + "d", dec) + sp \
               + struct.pack(order + "h", mag) + struct.pack(order + "f", pra) + struct.pack(order + "f", pdec)
    assert len(out) == 28 + 32 * len(STARS)
    return out


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    for name, order in (("le", "<"), ("be", ">")):
        with open(os.path.join(here, f"bsc5_three_stars.{name}.bin"), "wb") as f:
            f.write(build(order))
    print("wrote bsc5_three_stars.{le,be}.bin")

"""Synthetic star catalog + sky grid (docs/MODEL.md §8).

The reference points at the Yale Bright Star Catalog (`utils/BSC5PKL.pkl`, /root/reference/constants.py:32-33) but
the file is absent from the mount, and BASELINE.json's north_star asks for a *synthetic* catalog "binned on a sky
grid".  This module is product code: it authors the catalog (once; the result is committed as an .npz so nothing
depends on numpy's generator stream) and builds the declination-band x RA-cell grid the CUDA kernels stage into
shared memory.  The grid is an acceleration structure only; detection never depends on it.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

CATALOG_SEED = 20231105
CATALOG_SIZE = 9110
MAG_MIN, MAG_MAX, MAG_SLOPE = -1.5, 7.0, 0.5
DEFAULT_CELL_DEG = 3.0

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_CATALOG_PATH = os.path.join(_HERE, "utils", "catalog_synth_9110.npz")


def generate_catalog(n: int = CATALOG_SIZE, seed: int = CATALOG_SEED):
    """Uniform directions on S^2, magnitudes with dlogN/dm = 0.5 on [-1.5, 7]. Returns (xyz f64 [n,3], mag f32 [n])."""
    rng = np.random.default_rng(seed)
    z = rng.uniform(-1.0, 1.0, n)
    lon = rng.uniform(0.0, 2.0 * np.pi, n)
    rxy = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    xyz = np.stack([rxy * np.cos(lon), rxy * np.sin(lon), z], axis=1)
    xyz /= np.linalg.norm(xyz, axis=1, keepdims=True)
    lo, hi = 10.0 ** (MAG_SLOPE * MAG_MIN), 10.0 ** (MAG_SLOPE * MAG_MAX)
    mag = np.log10(lo + rng.uniform(0.0, 1.0, n) * (hi - lo)) / MAG_SLOPE
    return np.ascontiguousarray(xyz, dtype=np.float64), mag.astype(np.float32)


def save_catalog(path: str = DEFAULT_CATALOG_PATH, n: int = CATALOG_SIZE, seed: int = CATALOG_SEED) -> None:
    """The committed file is stored in sky-grid order (DEFAULT_CELL_DEG), so star_id == row index == position in the
    sorted device catalog: the kernels then take the Philox star key straight from their candidate lists."""
    xyz, mag = generate_catalog(n, seed)
    g = build_sky_grid(xyz, mag, DEFAULT_CELL_DEG)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    np.savez(path, xyz=g.xyz, mag=g.mag, seed=np.int64(seed))


def load_catalog(path: str | None = None):
    """Returns (xyz f64 [S,3], mag f32 [S]); star_id == row index."""
    with np.load(path or DEFAULT_CATALOG_PATH) as f:
        return np.ascontiguousarray(f["xyz"], dtype=np.float64), np.ascontiguousarray(f["mag"], dtype=np.float32)


@dataclass
class SkyGrid:
    """Stars sorted by (declination band, RA cell).

    Band b covers dec in [-pi/2 + b*band_h, -pi/2 + (b+1)*band_h); it has n_ra[b] equal RA cells over [0, 2pi).
    Flat cell index = cell_base[b] + ra_cell; stars of flat cell c are rows cell_start[c]:cell_start[c+1] of the
    sorted arrays.
    """
    n_bands: int
    band_h: float
    n_ra: np.ndarray        # int32 [n_bands]
    cell_base: np.ndarray   # int32 [n_bands + 1]
    cell_start: np.ndarray  # int32 [n_cells + 1]
    xyz: np.ndarray         # f64 [S,3]  sorted
    mag: np.ndarray         # f32 [S]    sorted
    star_id: np.ndarray     # int32 [S]  row in the catalog file (Philox key, MODEL.md §5)

    @property
    def n_cells(self) -> int:
        return int(self.cell_base[-1])

    @property
    def n_stars(self) -> int:
        return int(self.xyz.shape[0])


def build_sky_grid(xyz: np.ndarray, mag: np.ndarray, cell_deg: float = DEFAULT_CELL_DEG) -> SkyGrid:
    xyz = np.ascontiguousarray(xyz, dtype=np.float64)
    mag = np.ascontiguousarray(mag, dtype=np.float32)
    S = xyz.shape[0]
    n_bands = int(round(180.0 / cell_deg))
    band_h = np.pi / n_bands
    dec = np.arcsin(np.clip(xyz[:, 2], -1.0, 1.0))
    ra = np.mod(np.arctan2(xyz[:, 1], xyz[:, 0]), 2.0 * np.pi)
    band = np.clip(np.floor((dec + 0.5 * np.pi) / band_h).astype(np.int64), 0, n_bands - 1)
    dec_mid = -0.5 * np.pi + (np.arange(n_bands) + 0.5) * band_h
    n_ra = np.maximum(1, np.round(2.0 * np.pi * np.cos(dec_mid) / band_h)).astype(np.int32)
    cell_base = np.zeros(n_bands + 1, dtype=np.int32)
    cell_base[1:] = np.cumsum(n_ra)
    ra_cell = np.minimum((ra / (2.0 * np.pi) * n_ra[band]).astype(np.int64), n_ra[band] - 1)
    flat = cell_base[band] + ra_cell
    order = np.argsort(flat, kind="stable")
    counts = np.bincount(flat, minlength=int(cell_base[-1]))
    cell_start = np.zeros(int(cell_base[-1]) + 1, dtype=np.int32)
    cell_start[1:] = np.cumsum(counts)
    assert cell_start[-1] == S
    return SkyGrid(n_bands=n_bands, band_h=float(band_h), n_ra=n_ra, cell_base=cell_base, cell_start=cell_start,
                   xyz=np.ascontiguousarray(xyz[order]), mag=np.ascontiguousarray(mag[order]),
                   star_id=order.astype(np.int32))


def load_bsc5(path: str):
    """Yale Bright Star Catalog, 5th ed., binary distribution (`BSC5`, /root/reference/constants.py:33 `BSC5BIN`; the file is
    absent from the reference mount — SURVEY.md §8f rank 4: "real-catalog loader if a catalog file is ever supplied").

    Format (28-byte header of seven int32: STAR0, STAR1, STARN, STNUM, MPROP, NMAG, NBENT; |STARN| entries of NBENT = 32
    bytes: XNO f32, SRA0 f64 [rad], SDEC0 f64 [rad], IS 2 chars, MAG i16 [V x 100], XRPM f32, XDPM f32 [rad/yr]); STARN < 0
    marks J2000 coordinates.  Either byte order is accepted.  Entries without a position (the catalog's 14 non-stellar
    rows: RA = Dec = 0) are dropped.  Returns (xyz f64 [S, 3], mag f32 [S], hr_number i32 [S]) — pass (xyz, mag) to
    `build_sky_grid` / `TrialEngine(catalog=...)`."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 28:
        raise ValueError(f"{path}: too short for a BSC5 header")
    for order in ("<", ">"):
        hdr = raw[:28].view(np.dtype(order + "i4"))
        if hdr[6] == 32 and 0 < abs(int(hdr[2])) <= (raw.size - 28) // 32:
            break
    else:
        raise ValueError(f"{path}: not a BSC5 binary catalog (NBENT != 32 in either byte order)")
    n = abs(int(hdr[2]))
    rec = np.dtype([("xno", order + "f4"), ("ra", order + "f8"), ("dec", order + "f8"), ("sp", "S2"), ("mag", order + "i2"),
                    ("pm_ra", order + "f4"), ("pm_dec", order + "f4")])
    assert rec.itemsize == 32
    ent = raw[28:28 + 32 * n].view(rec)
    keep = ~((ent["ra"] == 0.0) & (ent["dec"] == 0.0))
    ra, dec = ent["ra"][keep].astype(np.float64), ent["dec"][keep].astype(np.float64)
    xyz = np.stack([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)], axis=1)
    return (np.ascontiguousarray(xyz), (ent["mag"][keep].astype(np.float32) / np.float32(100.0)),
            ent["xno"][keep].astype(np.int32))


if __name__ == "__main__":  # python -m startrackermpm_b200.catalog  -> (re)writes the committed catalog
    save_catalog()
    x, m = load_catalog()
    print("wrote", DEFAULT_CATALOG_PATH, x.shape, m.shape, "mag<=5.5:", int((m <= 5.5).sum()))

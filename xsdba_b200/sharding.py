"""Static partition of the grid across GPUs and across HBM-sized slabs.

Gridpoints are independent in EQM / DQM / QDM (the reference only parallelises over such dims,
base.py:640-648), so rank ``r`` of ``world`` owns the contiguous lat band
``[floor(n_lat*r/world), floor(n_lat*(r+1)/world))`` of the (time, lat, lon) arrays and no collective
is needed on the data path."""
from __future__ import annotations


def lat_band(n_lat: int, rank: int, world: int) -> tuple[int, int]:
    return (n_lat * rank) // world, (n_lat * (rank + 1)) // world


def slabs(n_rows: int, rows_per_slab: int) -> list[tuple[int, int]]:
    """(first_row, n_rows) pieces that are streamed through HBM one after the other."""
    return [(r0, min(rows_per_slab, n_rows - r0)) for r0 in range(0, n_rows, rows_per_slab)]

"""Host-side partitioning of the PIC cycle over the GPUs of one box.

Two axes, one exchange each:

* **particles** -- independent given the grid fields: rank ``r`` owns an even share; every rank
  deposits into its own full-size ``J`` / ``Rho`` and the grids are summed (all-reduce, NCCL);
* **kx rows of the spectral state** -- after the x-FFT every kx row is independent through the DHT,
  the mode-coupling contractions, the Poisson correction and the PSATD advance, *except* for the
  real solver's mirror term ``-conj(f(-kx))`` (reference f90/fb_math.f90:35-36), which couples row
  ``i`` to row ``(Nx - i) mod Nx``.  Slabs are therefore made of mirror pairs.

Slab layout (``Nx`` divisible by ``2 * world``; ``L = Nx / world`` rows per rank)::

    rank 0   : [0, L/2)  +  {Nx/2}  +  [Nx - L/2 + 1, Nx)          (rows 0 and Nx/2 are self-mirrored)
    rank r>0 : [r L/2, (r+1) L/2)  +  [Nx - (r+1) L/2 + 1, Nx - r L/2]

Rows are stored in ascending global order.  With that order the mirror partner of local row ``j`` is
``(L - j - shift) mod L`` with ``shift = 0`` on rank 0 (exactly the single-GPU formula) and ``1``
elsewhere -- the spectral kernels take ``shift`` as a parameter and are otherwise unchanged.
"""
from __future__ import annotations

import numpy as np


def particle_range(n: int, rank: int, world: int):
    """[start, stop) of the even particle split (the remainder goes to the lowest ranks)."""
    base, rem = divmod(int(n), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def slab_supported(nx: int, world: int) -> bool:
    return world >= 1 and nx % (2 * world) == 0 and nx // world >= 2


def kx_slab_rows(nx: int, rank: int, world: int) -> np.ndarray:
    """Global kx-row indices owned by ``rank``, ascending (see the module docstring)."""
    if world == 1:
        return np.arange(nx, dtype=np.int64)
    if not slab_supported(nx, world):
        raise ValueError("Nx=%d is not divisible by 2*world=%d" % (nx, 2 * world))
    L = nx // world
    h = L // 2
    if rank == 0:
        rows = np.concatenate((np.arange(0, h), [nx // 2], np.arange(nx - h + 1, nx)))
    else:
        rows = np.concatenate((np.arange(rank * h, (rank + 1) * h), np.arange(nx - (rank + 1) * h + 1, nx - rank * h + 1)))
    return rows.astype(np.int64)


def mirror_shift(rank: int, world: int) -> int:
    """``shift`` of the local mirror map ``j -> (L - j - shift) mod L``."""
    return 0 if (world == 1 or rank == 0) else 1


def local_mirror(nrows: int, shift: int) -> np.ndarray:
    return (nrows - np.arange(nrows) - shift) % nrows


def slab_of(arr: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """Rows ``rows`` of a spectral array (kx is the first, fastest axis), Fortran-ordered."""
    return np.asfortranarray(arr[rows])

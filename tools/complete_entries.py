"""How many CSR entries of the P1 stiffness matrix receive ALL their element contributions from ONE tile of 448
Morton-ordered cells?  (Those could be written with a plain bulk store instead of zero-fill + reduce-add; the answer
decides whether splitting every tile's runs into "complete" and "shared" segments can pay -- see DESIGN.md §4.1.)

    python tools/complete_entries.py [n]        # unit cube, n^3 cubes of 6 Kuhn tetrahedra, default 64
"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/", 2)[0])
from dolfinx_mpc_b200 import generators as gen  # noqa: E402


def spread(v):
    v = v.astype(np.uint64) & np.uint64(0xFFFFF)
    out = np.zeros_like(v)
    for b in range(20):
        out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
    return out


def main(n=64, tile=448):
    mesh = gen.create_unit_cube(n, n, n)
    cells = np.asarray(mesh.x_dofmap if hasattr(mesh, "x_dofmap") else mesh.cells, dtype=np.int64)
    x = np.asarray(mesh.x)[:, :3]
    cen = x[cells].mean(axis=1)
    q = np.minimum((cen * (1 << 20)).astype(np.int64), (1 << 20) - 1)
    code = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    order = np.argsort(code, kind="stable")
    tile_of = np.empty(len(cells), dtype=np.int64)
    tile_of[order] = np.arange(len(cells)) // tile
    nn = x.shape[0]
    i, j = np.meshgrid(range(4), range(4), indexing="ij")
    key = (cells[:, i.ravel()] * nn + cells[:, j.ravel()]).ravel()  # (row, col) of every element entry
    tl = np.repeat(tile_of, 16)
    o = np.lexsort((tl, key))
    key, tl = key[o], tl[o]
    first = np.r_[True, key[1:] != key[:-1]]
    newpair = first | np.r_[True, tl[1:] != tl[:-1]]
    ent = np.cumsum(first) - 1
    tiles_per_entry = np.bincount(ent, weights=newpair).astype(np.int64)
    nnz = len(tiles_per_entry)
    complete = int((tiles_per_entry == 1).sum())
    # runs of complete / shared entries in CSR order (every switch is one more bulk operation per tile at least)
    c = tiles_per_entry == 1
    switches = int((c[1:] != c[:-1]).sum())
    print(f"n={n}: cells={len(cells)} tiles={int(tile_of.max()) + 1} nnz={nnz} complete={complete} "
          f"({complete / nnz:.3f}) mean tiles/entry={tiles_per_entry.mean():.2f} complete<->shared switches in CSR order={switches} "
          f"({switches / (int(tile_of.max()) + 1):.0f} per tile)")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 64)

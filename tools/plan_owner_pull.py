"""Design check (CPU, numpy) for round-2 item 1 of DESIGN.md §8: "owner pulls" gather-scatter inside the axhelm
launch.  For a box of n^3 elements at order N it builds an element processing order, assigns every gather row to
its LAST sharing element in that order (the owner), and reports what the kernel design has to be sized for:

  * lag: over all rows, the smallest distance (in CTA iterations = list positions // nCTA) between the owner
    and its other sharers -- the slack available for publishing and observing the sharers' completion flags;
  * per element: owned rows, neighbour values to pull, distinct neighbour elements (shared-memory table sizes).

Orders compared: natural (lexicographic) and 8-colour (parity of the element coordinates, classes in sequence,
lexicographic inside a class).
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nekrs_b200 import meshgen  # noqa: E402


def orders(n):
    e = np.arange(n ** 3)
    ix, iy, iz = e % n, (e // n) % n, e // (n * n)
    colour = (ix % 2) + 2 * (iy % 2) + 4 * (iz % 2)
    return {"natural": e, "8-colour": np.lexsort((e, colour))}


def analyse(N, n, order, nCTA=148):
    m = meshgen.box_mesh(N, (n, n, n))
    ids = m.global_ids
    Np = m.Np
    pos = np.empty(n ** 3, dtype=np.int64)
    pos[order] = np.arange(n ** 3)
    # rows: sort nodes by id
    o = np.argsort(ids, kind="stable")
    sid = ids[o]
    starts = np.flatnonzero(np.r_[True, sid[1:] != sid[:-1]])
    ends = np.r_[starts[1:], sid.size]
    cnt = ends - starts
    multi = cnt > 1
    elem_of = o // Np
    owned_rows = np.zeros(n ** 3, dtype=np.int64)
    pulled = np.zeros(n ** 3, dtype=np.int64)
    nbrs = [set() for _ in range(n ** 3)]
    min_lag_iter, min_lag_pos = 10 ** 9, 10 ** 9
    for s, e_ in zip(starts[multi], ends[multi]):
        el = elem_of[s:e_]
        p = pos[el]
        k = np.argmax(p)
        owner = el[k]
        owned_rows[owner] += 1
        pulled[owner] += el.size - 1
        for q in np.delete(el, k):
            nbrs[owner].add(int(q))
        others = np.delete(p, k)
        min_lag_pos = min(min_lag_pos, int(p[k] - others.max()))
        min_lag_iter = min(min_lag_iter, int(p[k] // nCTA - others.max() // nCTA))
    nn = np.array([len(s_) for s_ in nbrs])
    return dict(min_lag_positions=min_lag_pos, min_lag_cta_iterations=min_lag_iter,
                max_owned_rows=int(owned_rows.max()), max_pulled_values=int(pulled.max()),
                max_neighbour_elements=int(nn.max()),
                smem_bytes_per_element_fp64=int(pulled.max() * 8 + pulled.max() * 2 + owned_rows.max() * 2 + nn.max() * 4))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=7)
    ap.add_argument("--n", type=int, default=16)
    args = ap.parse_args()
    for name, order in orders(args.n).items():
        print(name, analyse(args.N, args.n, order))


if __name__ == "__main__":
    main()

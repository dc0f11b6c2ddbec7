"""Mirror of Bijection.hs: dense renumbering in first-occurrence order (Bijection.hs:16-32).
Pure host bookkeeping in the reference too (Set/Map inserts); it defines the canonical vertex ids
that make the GPU's min-index component labels canonical."""
from __future__ import annotations


def biject(xs):
    """-> (indexOf: dict a -> Int, aOfIndex: list).  Ints are assigned in input order, first occurrence wins."""
    index_of, a_of_index = {}, []
    for x in xs:
        if x not in index_of:
            index_of[x] = len(a_of_index)
            a_of_index.append(x)
    return index_of, a_of_index

"""Genotype container and the alpha -> discrete cell decoder (reference: genotype.py:6-45).

Pure host code on 9xK float matrices; results must match the reference bit for bit including
its tie-breaking (tuples (weight, name, edge) sorted ascending, last two kept)."""
from collections import namedtuple

import numpy as np

from .prim_ops import DownOps, UpOps, NormOps

Genotype = namedtuple('Genotype', ['down', 'up'])

# Canonical benchmark genotype "G0" (the reference ships none - its
# log/best_genotype.pkl is a run artefact).  1 242 760 parameters at the shipped widths.
G0 = Genotype(
    down=[('down_conv', 0), ('down_conv', 1), ('conv', 2), ('down_dep_conv', 1), ('dil_conv', 3),
          ('se_conv', 2)],
    up=[('up_conv', 1), ('conv', 0), ('up_dep_conv', 1), ('dil_conv', 2), ('up_dil_conv', 1),
        ('se_conv', 3)])


class GenoParser:
    def __init__(self, n_nodes):
        self.n_nodes = n_nodes

    def parse(self, alpha1, alpha2, downward=True):
        '''
        alpha1 / alpha2: softmaxed weights of the stride-1 / stride-2 MixedOps, one row per edge.
        Every edge keeps its strongest op; every node keeps its two strongest edges, where
        stride-2 edges are compared after rescaling by (#ops of their family / #NormOps).
        '''
        strided_family = DownOps if downward else UpOps
        picked = []
        row = 0
        for n_edges in range(2, 2 + self.n_nodes):
            cands = []
            for src in range(n_edges):
                strided = (src < 2) if downward else (src == 1)
                if strided:
                    k = np.argmax(alpha2[row])
                    # same expression order as the reference: (w * len(family)) / len(NormOps)
                    cands.append((alpha2[row][k] * len(strided_family) / len(NormOps),
                                  strided_family[k], src))
                else:
                    k = np.argmax(alpha1[row])
                    cands.append((alpha1[row][k], NormOps[k], src))
                row += 1
            cands.sort()
            picked += [(name, src) for _, name, src in cands[-2:]]
        return picked

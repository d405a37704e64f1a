"""Drop-in for the reference's DistIt (simulation_utilities/tensorflow_descriptors/distance_descriptors.py:6-213): the
distance, Coulomb or SPF descriptor of every walker, optionally with atoms sorted inside sub-lists and whole groups
swapped, as the upper triangle or the full matrix.  Same constructor, same errors, same numbers (bit for bit); the
work is done by the CUDA kernel behind pvd_distit, there is no NumPy/CuPy fallback."""
import itertools as itt

import numpy as np

__all__ = ['DistIt']

_METHODS = {'distance': 0, 'coulomb': 1, 'spf': 2}


class DistIt:
    def __init__(self, zs, method, eq_xyz=None, sorted_groups=None, sorted_atoms=None, full_mat=False, force_numpy=False):
        """zs: nuclear charges; method: 'spf', 'distance' or 'coulomb'; eq_xyz: equilibrium structure (spf);
        sorted_groups: equally long lists of atoms that may trade places as wholes; sorted_atoms: lists covering ALL
        atoms, each sorted within itself; full_mat: full matrix instead of the upper triangle (diagonal excluded).
        force_numpy is accepted for compatibility and ignored (reference :19, 38-46 choose between NumPy and CuPy)."""
        self.zs = np.asarray(zs)
        self.method = method.lower()
        self.eq_xyz = eq_xyz
        self.sorted_groups = sorted_groups
        self.sorted_atoms = sorted_atoms
        self.full_mat = full_mat
        self.force_numpy = force_numpy
        self._initialize()

    def _initialize(self):                      # reference :48-86
        if self.method not in _METHODS:
            raise ValueError("method must be 'spf', 'distance' or 'coulomb'")
        self.num_atoms = len(self.zs)
        self.sort_mat = self.sorted_atoms is not None or self.sorted_groups is not None
        if self.sorted_atoms is not None:
            if sum(len(lst) for lst in self.sorted_atoms) != self.num_atoms:
                raise ValueError("Please put all atoms in sorted_atoms list")
        if self.sorted_groups is not None:
            self.sorted_groups = np.asarray(self.sorted_groups)
        self.idxs = list(itt.combinations(range(self.num_atoms), 2))
        self.idxs_0 = [p[0] for p in self.idxs]
        self.idxs_1 = [p[1] for p in self.idxs]
        if self.method == 'coulomb':            # reference :155-165: Z_i Z_j off the diagonal, 0.5 Z^2.4 on it
            rest = np.ones((self.num_atoms, self.num_atoms))
            np.fill_diagonal(rest, 0.5 * self.zs ** 0.4)
            skeleton = np.outer(self.zs, self.zs) * rest
            self.diag_coulomb = np.ascontiguousarray(np.diag(skeleton), dtype=np.float64)
            self._pair_scale = np.ascontiguousarray(skeleton[self.idxs_0, self.idxs_1], dtype=np.float64)
        else:
            self.diag_coulomb = self._pair_scale = None
        self.r_eq = None
        if self.eq_xyz is None and self.method == 'spf':
            raise ValueError("eq_xyz is not set but using spf. Fix!")
        if self.method == 'spf':
            # reference :74-86: r_eq is the equilibrium structure's own distance descriptor, sorted the way the walkers' will be
            # two copies, as in the reference (:75): with a single walker NumPy would add the group totals in another order
            eq = np.repeat(np.asarray(self.eq_xyz, dtype=np.float64)[None], 2, axis=0)
            self.r_eq = self._launch(eq, 'distance', full_mat=self.sort_mat, r_eq=None)[0]

    def _launch(self, cds, method, full_mat, r_eq):
        from ..._capi import check, f64, lib, ptr
        cds = f64(cds)
        if cds.ndim != 3 or cds.shape[1] != self.num_atoms or cds.shape[2] != 3:
            raise ValueError(f"expected coordinates of shape (n, {self.num_atoms}, 3)")
        n, na = cds.shape[0], self.num_atoms
        out = np.empty((n, na, na) if full_mat else (n, len(self.idxs)))
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)                       # noqa: E731
        lists = ofs = groups = None
        nl = ng = gs = 0
        if self.sorted_atoms is not None:
            lists = i32([a for lst in self.sorted_atoms for a in lst])
            ofs = i32(np.concatenate([[0], np.cumsum([len(lst) for lst in self.sorted_atoms])]))
            nl = len(self.sorted_atoms)
        if self.sorted_groups is not None:
            groups, (ng, gs) = i32(self.sorted_groups.ravel()), self.sorted_groups.shape
        req = None if r_eq is None else f64(r_eq)
        null = lambda a: None if a is None else ptr(a)                                # noqa: E731
        check(lib.pvd_distit(ptr(cds), n, na, _METHODS[method], null(self._pair_scale if method == 'coulomb' else None),
                             null(self.diag_coulomb if method == 'coulomb' else None), null(req), null(lists), null(ofs), nl,
                             null(groups), int(ng), int(gs), 1 if full_mat else 0, ptr(out)))
        return out

    def run(self, cds):
        """Cartesian coordinates (n, num_atoms, 3) -> descriptor, vector (upper triangle) or matrix form (reference :177-213)."""
        return self._launch(np.asarray(cds), self.method, self.full_mat, self.r_eq)

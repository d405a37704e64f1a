"""Test-only stand-in for h5py so that the *reference* package imports in this image
(h5py is not installed; reference import sites: pyvibdmc/analysis/extract_sim_info.py:1,
pyvibdmc/simulation_utilities/sim_archive.py:1).  Used only by tests/golden/make_golden.py
when generating fixtures from /root/reference.  Writes datasets into an .npz next to the
requested file name; never used by the product."""
import numpy as np


class File:
    def __init__(self, fname, mode="r"):
        self.fname, self.mode, self._d = fname, mode, {}
        if mode == "r":
            self._d = dict(np.load(fname + ".npz"))

    def create_dataset(self, key, data=None):
        self._d[key] = np.asarray(data)

    def __getitem__(self, k):
        return self._d[k]

    def __enter__(self):
        return self

    def __exit__(self, *a):
        if self.mode != "r":
            np.savez(self.fname + ".npz", **self._d)
        return False
